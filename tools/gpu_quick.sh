#!/bin/bash
# tests + bench + graph trace; usage: bash tools/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-rXX}; K=${2:-}
O=gpurun_out; mkdir -p $O
if [ -n "$K" ]; then timeout 1500 python -m pytest tests -m gpu -q -x -k "$K" > $O/${TAG}_pytest.log 2>&1; else timeout 1700 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; fi
echo "pytest rc=$?"; grep -n "^E  \|passed\|failed\|^FAILED\|^ERROR" $O/${TAG}_pytest.log | head -30
timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench rc=$?"
grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench_n1.json | head -3; tail -3 $O/${TAG}_bench_n1.err
timeout 300 python tools/graph_trace.py $O/${TAG}_graph_trace.txt > $O/${TAG}_graph_trace.log 2>&1; echo "trace rc=$?"; head -1 $O/${TAG}_graph_trace.txt
