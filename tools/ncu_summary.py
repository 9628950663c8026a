"""Write a text summary of an ncu --set full report into profiles/: key metrics, stall reasons, pipe utilisation,
opcode mix per pixel.  usage: python tools/ncu_summary.py report.ncu-rep kernel_regex pixels_per_launch out.txt [note]"""
import collections
import csv
import re
import subprocess
import sys

rep, rx, px, out = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
note = sys.argv[5] if len(sys.argv) > 5 else ''
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '--kernel-name', 'regex:' + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
import os
r = data[int(os.environ.get('NCU_INDEX', '0'))]
lines = []
w = lines.append
w('ncu --set full --clock-control none  (report: %s, %d launch(es) of the kernel captured; first one shown)' % (rep.split('/')[-1], len(data)))
if note:
    w(note)
w('kernel: ' + r[idx['Kernel Name']])
w('grid %s  block %s  pixels per launch %.0f' % (r[idx['Grid Size']], r[idx['Block Size']], px))
w('')
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.avg.per_second',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active']
for k in keys:
    if k in idx:
        w('%-78s %18s %s' % (k, r[idx[k]], units[idx[k]]))
try:
    inst = float(r[idx['smsp__inst_executed.sum']].replace(',', ''))
    w('%-78s %18.1f' % ('thread instructions per pixel', inst * 32 / px))
    rd = float(r[idx['dram__bytes_read.sum']]); wr = float(r[idx['dram__bytes_write.sum']])
    mult = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}
    tb = rd * mult[units[idx['dram__bytes_read.sum']]] + wr * mult[units[idx['dram__bytes_write.sum']]]
    w('%-78s %18.2f B/px  (%.0f bytes per launch)' % ('DRAM traffic (read + write)', tb / px, tb))
except Exception as e:
    w('(derived metrics failed: %s)' % e)
w('')
w('warp stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active.ratio)')
st = []
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and 'per_issue_active' in h:
        try:
            st.append((float(r[i]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
        except ValueError:
            pass
for v, n in sorted(st, reverse=True):
    if v > 0.02:
        w('    %-28s %6.2f' % (n, v))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx], capture_output=True, text=True).stdout
ops = collections.Counter()
tot = 0
sidx = None
nk = 0
for row in csv.reader(src.splitlines()):
    if row and row[0] == 'Kernel Name':
        nk += 1
        if nk > 1:
            break
    if row and row[0] == 'Address':
        sidx = {h: i for i, h in enumerate(row)}
        continue
    if sidx is None or len(row) < 8 or not row[0].startswith('0x'):
        continue
    try:
        n = int(row[sidx['Thread Instructions Executed']])
    except ValueError:
        continue
    s = row[sidx['Source']].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', s)
    op = m.group(2) if m else s[:12]
    op = '.'.join(op.split('.')[:2]) if op.startswith(('LDS', 'STS', 'LDG', 'STG', 'F2F', 'I2F', 'MUFU', 'UTMA', 'SYNCS')) else op.split('.')[0]
    ops[op] += n
    tot += n
w('')
w('opcode mix, thread instructions per pixel (first captured launch)')
for op, n in ops.most_common(32):
    w('    %-16s %7.2f   %5.1f%%' % (op, n / px, 100.0 * n / max(tot, 1)))
open(out, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines[:40]))
