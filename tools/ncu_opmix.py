"""Opcode mix (thread instructions per pixel) of one kernel from an ncu report's source page.
usage: python tools/ncu_opmix.py report.ncu-rep kernel_regex [pixels_per_launch] [top]"""
import collections
import csv
import re
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
px = float(sys.argv[3]) if len(sys.argv) > 3 else 3000 * 4096
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '--kernel-name', 'regex:' + rx], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
nl = len(rr) - 2
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
ops = collections.Counter()
tot = 0
idx = None
for r in rows:
    if r and r[0] == 'Address':
        idx = {h: i for i, h in enumerate(r)}
        continue
    if idx is None or len(r) < 8 or not r[0].startswith('0x'):
        continue
    try:
        n = int(r[idx['Thread Instructions Executed']])
    except ValueError:
        continue
    src = r[idx['Source']].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    op = m.group(2) if m else src[:12]
    op = '.'.join(op.split('.')[:2]) if op.startswith(('LDS', 'STS', 'LDG', 'STG', 'F2F', 'I2F', 'MUFU')) else op.split('.')[0]
    ops[op] += n
    tot += n
print('launches in report: %d   thread-instr/px: %.1f' % (nl, tot / nl / px))
for op, n in ops.most_common(top):
    print('%-14s %7.2f/px  %5.1f%%' % (op, n / nl / px, 100.0 * n / tot))
