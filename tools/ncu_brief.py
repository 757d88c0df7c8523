"""Brief of an .ncu-rep: per kernel the headline metrics, and for one kernel the stall reasons / opcode mix of its
source page.  `python tools/ncu_brief.py <report> [kernel regex]` (development tool)."""
import csv, subprocess, sys, io
from collections import defaultdict
rep = sys.argv[1]
rx = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct', 'sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed']
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print('----', r[idx['Kernel Name']][:70])
    for w in want:
        if w in idx:
            print(f"  {w} = {r[idx[w]]} {units[idx[w]]}")
if rx:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    tot = {}
    for i, h in enumerate(hdr):
        if h.startswith('stall_') and 'Not' not in h:
            s = 0
            for r in rows[2:]:
                try: s += int(r[i])
                except Exception: pass
            tot[h] = s
    T = sum(tot.values()) or 1
    print('stalls:', {k: f"{v / T * 100:.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]})
    i_src, i_ex = hdr.index('Source'), hdr.index('Instructions Executed')
    g = defaultdict(int)
    n = 0
    for r in rows[2:]:
        try: ex = int(r[i_ex])
        except Exception: continue
        t = r[i_src].split()
        op = t[1] if t[0].startswith('@') else t[0]
        g['.'.join(op.split('.')[:2])] += ex
        n += ex
    print('warp instructions', n)
    for op, v in sorted(g.items(), key=lambda kv: -kv[1])[:14]:
        print(f"  {op:12s} {v:>12d} {v / n * 100:5.1f}%")
