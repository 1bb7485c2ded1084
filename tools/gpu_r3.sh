#!/bin/bash
# round-2 iteration visit: parity subset, R-GCN / sweep micro-benchmarks, bench line, graph trace
# usage (under gpurun): bash tools/gpu_r3.sh <tag> [pytest -k expr]
TAG=${1:-rXX}; K=${2:-"rgcn or sweep or benched or tip_model or dd_net"}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "$K" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log
timeout 200 python tools/ubench_rgcn.py > $O/${TAG}_ubench_rgcn.json 2> $O/${TAG}_ubench_rgcn.err; cat $O/${TAG}_ubench_rgcn.json
timeout 200 python tools/ubench_sweep.py > $O/${TAG}_sweep.json 2> $O/${TAG}_sweep.err; cat $O/${TAG}_sweep.json
timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench rc=$?"
grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench_n1.json | head -2; tail -2 $O/${TAG}_bench_n1.err
timeout 300 python tools/graph_trace.py $O/${TAG}_graph_trace.txt > $O/${TAG}_graph_trace.log 2>&1; echo "trace rc=$?"; head -1 $O/${TAG}_graph_trace.txt
