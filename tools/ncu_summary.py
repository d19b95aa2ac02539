"""Summarise an ncu report exported with `--page raw --csv` and `--page source --csv`.
usage: python tools/ncu_summary.py raw.csv [src.csv]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    g = lambda k: d.get(k, "")
    print("kernel:", g("Kernel Name")[:60], "grid", g("Grid Size"), "block", g("Block Size"))
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg", "smsp__warps_eligible.avg.per_cycle_active"]
    for k in keys:
        if k in d: print(f"  {k:72s} {d[k]:>16s} {units[hdr.index(k)]}")
    for k in hdr:
        if "average_warps_issue_stalled" in k and "not_issued" not in k:
            try:
                v = float(d[k])
            except ValueError:
                continue
            if v > 0.05: print(f"  stall {k.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {v:6.2f}")
if len(sys.argv) > 2:
    rows = list(csv.reader(open(sys.argv[2])))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    h = rows[hi]; ci, ce, cs = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    ops, st = collections.Counter(), collections.Counter(); tot = stot = 0
    for r in rows[hi + 1:]:
        if len(r) <= max(ce, cs): continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ci])
        if not m: continue
        try: n, s_ = int(r[ce]), int(r[cs])
        except ValueError: continue
        op = m.group(2).split(".")[0]; ops[op] += n; st[op] += s_; tot += n; stot += s_
    print("opcode mix (share of executed warp instructions | share of stall samples)")
    for op, n in ops.most_common(28): print(f"  {op:10s} {100*n/tot:5.1f}%  {100*st[op]/max(stot,1):5.1f}%")
