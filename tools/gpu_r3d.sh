#!/bin/bash
# stream-priority experiment: (main, side) = (0,-1) current, (0,0), (-1,0); bench line + graph trace each
TAG=${1:-rXX}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "rgcn or benched" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest.log
for V in "0 -1" "0 0" "-1 0"; do
  set -- $V; M=$1; S=$2
  TIPB_BENCH_MAIN_PRIORITY=$M TIPB_SIDE_PRIORITY=$S timeout 600 python bench.py --steps 40 --warmup 5 --skip-cpu-baseline > $O/${TAG}_bench_m${M}_s${S}.json 2> $O/${TAG}_bench_m${M}_s${S}.err; echo "main $M side $S rc=$?"
  grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench_m${M}_s${S}.json | head -2
  TIPB_BENCH_MAIN_PRIORITY=$M TIPB_SIDE_PRIORITY=$S timeout 300 python tools/graph_trace.py $O/${TAG}_graph_trace_m${M}_s${S}.txt > $O/${TAG}_graph_trace.log 2>&1; head -1 $O/${TAG}_graph_trace_m${M}_s${S}.txt
done
