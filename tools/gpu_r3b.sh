#!/bin/bash
# iteration visit: parity subset, R-GCN micro-benchmark, bench line with the sampler stream at both priorities, graph traces
TAG=${1:-rXX}; K=${2:-"rgcn or benched or tip_model or dd_net"}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "$K" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log
timeout 200 python tools/ubench_rgcn.py > $O/${TAG}_ubench_rgcn.json 2> $O/${TAG}_ubench_rgcn.err; cat $O/${TAG}_ubench_rgcn.json
python -c "from tip_b200 import _lib; print('tc status', _lib.lib().tipb_rgcn_tc_status())"
for P in -1 0; do
  TIPB_SIDE_PRIORITY=$P timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline > $O/${TAG}_bench_n1_p$P.json 2> $O/${TAG}_bench_n1_p$P.err; echo "bench prio $P rc=$?"
  grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench_n1_p$P.json | head -2
  TIPB_SIDE_PRIORITY=$P timeout 300 python tools/graph_trace.py $O/${TAG}_graph_trace_p$P.txt > $O/${TAG}_graph_trace.log 2>&1; head -1 $O/${TAG}_graph_trace_p$P.txt
done
