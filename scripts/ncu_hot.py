"""Summarise an .ncu-rep: headline metrics + hottest SASS instructions with their dominant stall reasons."""
import csv, subprocess, sys, io
rep = sys.argv[1]
sel = ["--launch-skip", sys.argv[3], "--launch-count", "1"] if len(sys.argv) > 3 else []  # which launch of the report
raw = subprocess.run(["ncu", "-i", rep] + sel + ["--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
for h, u, v in zip(hdr, unit, vals):
    if h in want: print(f"{h:75s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep] + sel + ["--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ci, cs = hdr.index("Source"), hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data, seen = [], set()
for r in rows[2:]:  # (the source page repeats every SASS row once per view: keep one)
    if len(r) > cs and r[cs].isdigit() and tuple(r) not in seen:
        seen.add(tuple(r))
        data.append((int(r[cs]), r))
tot = sum(d[0] for d in data)
agg = {hdr[i]: sum(int(r[i]) for _, r in data) for i in stall}
print("stall totals:", sorted(agg.items(), key=lambda x: -x[1])[:7])
for n, r in sorted(data, key=lambda x: -x[0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 22]:
    st = sorted([(int(r[i]), hdr[i][6:]) for i in stall], reverse=True)[:2]
    print(f"{n:6d} {100*n/tot:5.1f}%  {r[ci].strip()[:64]:64s} {st}")
