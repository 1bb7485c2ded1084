#!/bin/bash
# One GPU-box visit: parity tests, both bench arms, ncu launch list, a metrics pass over every library kernel of two
# eager steps and an `ncu --set full` capture of the top kernels (exported to CSV on the box: gpurun_out is capped at
# 64 MiB).  usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^E  \|passed\|failed\|^FAILED" $O/${TAG}_pytest.log | head
timeout 600 python bench.py --steps 30 --warmup 5 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench rc=$?"
cat $O/${TAG}_bench_n1.json
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/${TAG}_smoke.log
timeout 300 python bench.py --mod add --steps 30 --warmup 5 --skip-cpu-baseline > $O/${TAG}_bench_add_n1.json 2> $O/${TAG}_bench_add_n1.err; echo "TIP-add: $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench_add_n1.json | head -1)"
timeout 120 python tools/ubench_sweep.py > $O/${TAG}_sweep.json 2> $O/${TAG}_sweep.err; cat $O/${TAG}_sweep.json
if [ "$WITH_REFERENCE_ARM" = "1" ]; then
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; echo "reference arm rc=$?"
cat $O/${TAG}_bench_reference.json
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches_eager.csv \
    python tools/one_step.py 3 > $O/${TAG}_launches.log 2>&1; echo "launch list rc=$?"
read SKIP_ALL SKIP_TOP <<< $(python - <<PY
import csv, re
rows = [r for r in csv.reader(l for l in open("$O/${TAG}_launches_eager.csv") if l.startswith('"'))]
names = [r[4] for r in rows[1:]]
gen = [i for i, n in enumerate(names) if "k_mt_generate" in n]
start = gen[-3] + 1 if len(gen) >= 3 else 0          # two eager steps before the end
top = re.compile(r"k_seg_aggregate_flat|k_decoder_seg|k_grp_place")
start1 = gen[-2] + 1 if len(gen) >= 2 else 0
print(sum(1 for i, n in enumerate(names) if i < start and "tipb::" in n), sum(1 for i, n in enumerate(names) if i < start1 and top.search(n)))
PY
)
echo "skip $SKIP_ALL library launches (metrics pass), $SKIP_TOP top-kernel launches (full capture)"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__block_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,lts__t_sector_hit_rate.pct
TIPB_DUMP_WORKLOAD=$O/${TAG}_workload.json timeout 600 ncu --metrics $M --clock-control none -k 'regex:^k_' -s $SKIP_ALL -c 130 --csv --page raw \
    --log-file $O/${TAG}_allkernels_raw.csv python tools/one_step.py 3 > $O/${TAG}_all.log 2>&1; echo "metrics pass rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_seg_aggregate_flat|k_decoder_seg|k_grp_place' \
    -s $SKIP_TOP -c 7 -f -o /tmp/${TAG}_top python tools/one_step.py 3 > $O/${TAG}_top.log 2>&1; echo "full capture rc=$?"
ncu -i /tmp/${TAG}_top.ncu-rep --page raw --csv > $O/${TAG}_top_raw.csv 2>/dev/null
for k in k_seg_aggregate_flat k_decoder_seg k_grp_place; do
    ncu -i /tmp/${TAG}_top.ncu-rep --page source --csv --kernel-name regex:$k --launch-count 1 > $O/${TAG}_src_$k.csv 2>/dev/null
done
timeout 200 python tools/graph_trace.py $O/${TAG}_graph_trace.txt > $O/${TAG}_graph_trace.log 2>&1; echo "graph trace rc=$?"; head -1 $O/${TAG}_graph_trace.txt
du -sm $O
