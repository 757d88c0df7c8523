"""Summarise one kernel of an .ncu-rep: headline metrics + top stall sites (development tool)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2 + idx]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "sm__warps_active.avg.per_cycle_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for h, u, v in zip(hdr, units, vals):
    if h in want or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(f(r, "# Samples") for r in data)
print("total samples", tot, "instructions", len(data))
print({k: int(sum(f(r, k) for r in data)) for k in keys})
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    st = {k: f(r, k) for k in keys if f(r, k) > 0}
    main = max(st, key=st.get) if st else ""
    print(r[ix["Address"]][-5:], int(f(r, "# Samples")), int(f(r, "Instructions Executed")), main, r[ix["Source"]][:100])
