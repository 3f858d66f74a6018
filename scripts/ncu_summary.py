"""Summarise an ncu report: key metrics per kernel instance + top stall lines from the source page."""
import csv, subprocess, sys, io
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units, rows = r[0], r[1], r[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
idx = {h: i for i, h in enumerate(hdr)}
_seen_k = set()
for row in rows:
    kn = row[idx["Kernel Name"]][:48] + row[idx["launch__grid_size"]]
    if kn in _seen_k:
        continue
    _seen_k.add(kn)
    print("=" * 100)
    for w in want:
        if w in idx:
            v = row[idx[w]]
            print(f"  {w.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', ''):75s} {v[:90]} {units[idx[w]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdrs = [i for i, r_ in enumerate(rows) if r_ and r_[0] == "Address"]
names = [rows[i - 1][1] if i > 0 else "" for i in hdrs]
seen = set()
for k, hi in enumerate(hdrs):
    name = names[k][:60]
    if name in seen:
        continue
    seen.add(name)
    h = rows[hi]
    end = hdrs[k + 1] - 1 if k + 1 < len(hdrs) else len(rows)
    body = rows[hi + 1:end]
    si, so, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
    tot = sum(int(b[si]) for b in body if len(b) > si and b[si].isdigit())
    print("-" * 100)
    print(name, "total samples", tot)
    top = sorted([(int(b[si]), i, b[so].strip(), b[ie]) for i, b in enumerate(body) if len(b) > si and b[si].isdigit()], reverse=True)[:topn]
    for s_, i, t, e in top:
        print(f"  {s_:6d} {s_ / max(tot, 1) * 100:5.1f}%  #{i:5d} exec={e:>9s}  {t[:100]}")
