#!/bin/bash
# One single-GPU evidence visit (round 2): ncu launch list of the bench command, a metrics pass over every library kernel
# of one eager step (-> per-kernel roofline table, kernel_traffic.json), an `ncu --set full` capture of the largest kernels
# (raw + source pages exported to CSV on the box: gpurun_out is capped at 64 MiB), the decoder sweep micro-benchmark.
# usage (from the repo root, under gpurun): bash tools/gpu_round2.sh <tag>
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > $O/${TAG}_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches_eager.csv \
    python tools/one_step.py 3 > $O/${TAG}_launches.log 2>&1; echo "eager launch list rc=$?"
read SKIP_ALL SKIP_TOP <<< $(python - <<PY
import csv, re
rows = [r for r in csv.reader(l for l in open("$O/${TAG}_launches_eager.csv") if l.startswith('"'))]
names = [r[4] for r in rows[1:]]
gen = [i for i, n in enumerate(names) if "k_mt_generate" in n]
start = gen[-2] + 1 if len(gen) >= 2 else 0          # the last eager step
top = re.compile(r"k_pair_pass|k_rgcn_node|k_seg_aggregate_flat|k_window_scan|k_materialize_main")
print(sum(1 for i, n in enumerate(names) if i < start and "tipb::" in n), sum(1 for i, n in enumerate(names) if i < start and top.search(n)))
PY
)
echo "skip $SKIP_ALL library launches (metrics pass), $SKIP_TOP top-kernel launches (full capture)"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__block_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,lts__t_sector_hit_rate.pct
TIPB_DUMP_WORKLOAD=$O/${TAG}_workload.json timeout 600 ncu --metrics $M --clock-control none -k 'regex:^k_' -s $SKIP_ALL -c 80 --csv --page raw \
    --log-file $O/${TAG}_allkernels_raw.csv python tools/one_step.py 3 > $O/${TAG}_all.log 2>&1; echo "metrics pass rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_pair_pass|k_rgcn_node|k_seg_aggregate_flat|k_window_scan|k_materialize_main' \
    -s $SKIP_TOP -c 11 -f -o /tmp/${TAG}_top python tools/one_step.py 3 > $O/${TAG}_top.log 2>&1; echo "full capture rc=$?"
ncu -i /tmp/${TAG}_top.ncu-rep --page raw --csv > $O/${TAG}_top_raw.csv 2>/dev/null
for k in k_pair_pass k_rgcn_node_fwd_tc k_rgcn_node_bwd_tc k_seg_aggregate_flat; do
    ncu -i /tmp/${TAG}_top.ncu-rep --page source --csv --kernel-name regex:$k --launch-count 1 > $O/${TAG}_src_$k.csv 2>/dev/null
done
timeout 200 python tools/ubench_sweep.py > $O/${TAG}_sweep.json 2> $O/${TAG}_sweep.err; cat $O/${TAG}_sweep.json
du -sm $O
