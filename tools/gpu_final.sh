#!/bin/bash
# final evidence visit of a round (one GPU): every GPU test, smoke(), the bench lines of every BASELINE.json config that fits
# one GPU (default line WITH the CPU baseline), the reference arm, graph trace, then the ncu passes of tools/gpu_round2.sh
TAG=${1:-rXX}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $O/${TAG}_smi.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/${TAG}_smoke.log
timeout 900 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench rc=$?"; grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench_n1.json | head -2
timeout 600 python bench.py --mod add --skip-cpu-baseline > $O/${TAG}_bench_add_n1.json 2> $O/${TAG}_bench_add_n1.err; echo "bench add rc=$?"; grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench_add_n1.json | head -1
timeout 600 python bench.py --model dd --skip-cpu-baseline > $O/${TAG}_dd_n1.json 2> $O/${TAG}_dd_n1.err; echo "bench dd rc=$?"; grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_dd_n1.json | head -1
timeout 300 python bench.py --workload sweep > $O/${TAG}_bench_sweep.json 2> $O/${TAG}_bench_sweep.err; echo "bench sweep rc=$?"; grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench_sweep.json | head -1
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; echo "reference arm rc=$?"; cut -c1-300 $O/${TAG}_bench_reference.json
TIPB_BENCH_MAIN_PRIORITY=-1 timeout 300 python tools/graph_trace.py $O/${TAG}_graph_trace.txt > $O/${TAG}_graph_trace.log 2>&1; head -1 $O/${TAG}_graph_trace.txt
bash tools/gpu_round2.sh $TAG
