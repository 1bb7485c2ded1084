"""summarise an `ncu --set full` capture brought back as CSV pages (tools/gpu_ncu_kernel.sh): headline metrics, stall
reasons, and instructions / stall samples per code segment between barriers.  usage: python tools/ncu_summary.py <tag>"""
import csv, sys
tag = sys.argv[1]
rows = list(csv.reader(open('gpurun_out/%s_raw.csv' % tag)))
hdr = rows[0]; col = {n: i for i, n in enumerate(hdr)}
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print(r[col["Kernel Name"]][:70])
    for k in want:
        if k in col: print("   %-80s %s %s" % (k, r[col[k]], rows[1][col[k]]))
    for k in col:
        if "stalled" in k and "per_issue_active" in k and "not_issued" not in k:
            v = float(r[col[k]] or 0)
            if v > 0.3: print("      stall %-40s %.2f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
rows = list(csv.reader(open('gpurun_out/%s_source.csv' % tag)))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
seen = set()
for ki, st in enumerate(starts):
    en = starts[ki + 1] if ki + 1 < len(starts) else len(rows)
    if rows[st][1] in seen: continue
    seen.add(rows[st][1])
    hdr = rows[st + 1]; col = {n: i for i, n in enumerate(hdr)}
    body = rows[st + 2:en]
    total = sum(int(r[col["Instructions Executed"]] or 0) for r in body); tsm = sum(int(r[col["# Samples"]] or 0) for r in body)
    segs = []; cur = [0, 0, 0]
    for r in body:
        cur[0] += int(r[col["Instructions Executed"]] or 0); cur[1] += int(r[col["# Samples"]] or 0); cur[2] += 1
        if "BAR.SYNC" in r[col["Source"]]: segs.append(tuple(cur)); cur = [0, 0, 0]
    segs.append(tuple(cur))
    print(rows[st][1][:60], "total inst", total, "samples", tsm)
    for i, s in enumerate(segs): print("  seg %d: inst %11d (%.1f%%) samples %.1f%%  lines %d" % (i, s[0], 100 * s[0] / max(total, 1), 100 * s[1] / max(tsm, 1), s[2]))
    if len(sys.argv) > 2:
        top = sorted(body, key=lambda r: -int(r[col["# Samples"]] or 0))[:int(sys.argv[2])]
        for r in top: print("     %6s samples %9s inst  %s" % (r[col["# Samples"]], r[col["Instructions Executed"]], r[col["Source"]][:90]))
