"""Summarise an ncu report: per kernel duration, DRAM bytes, executed instructions per pixel, pipe utilisation.
usage: python tools/ncu_brief.py report.ncu-rep [pixels_per_launch]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
px = float(sys.argv[2]) if len(sys.argv) > 2 else 3000 * 4096
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size',
        'sm__cycles_elapsed.avg.per_second', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'smsp__average_warp_latency_issue_stalled_barrier', 'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_not_selected_per_warp_active.pct', 'smsp__warp_issue_stalled_wait_per_warp_active.pct']
seen = {}
for r in data:
    name = r[idx['Kernel Name']]
    seen.setdefault(name, []).append(r)
for name, rs in seen.items():
    r = rs[len(rs) // 2]
    print('----', name[:110], ' (%d launches captured)' % len(rs))
    for w in want:
        if w in idx:
            print('    %-75s %s %s' % (w, r[idx[w]], units[idx[w]]))
    try:
        inst = float(r[idx['smsp__inst_executed.sum']].replace(',', ''))
        print('    thread-instructions per pixel (px=%g): %.1f' % (px, inst * 32 / px))
    except Exception:
        pass
