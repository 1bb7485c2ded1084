#!/bin/bash
# enqueue-order experiment: encoder first (default) vs sampler first; bench line + graph trace each
TAG=${1:-rXX}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "rgcn or benched or tip_model" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest.log
for V in 1 0; do
  TIPB_ENCODER_FIRST=$V timeout 600 python bench.py --steps 40 --warmup 5 --skip-cpu-baseline > $O/${TAG}_bench_ef$V.json 2> $O/${TAG}_bench_ef$V.err; echo "encoder_first $V rc=$?"
  grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench_ef$V.json | head -2
  TIPB_BENCH_MAIN_PRIORITY=-1 TIPB_ENCODER_FIRST=$V timeout 300 python tools/graph_trace.py $O/${TAG}_graph_trace_ef$V.txt > $O/${TAG}_graph_trace.log 2>&1; head -1 $O/${TAG}_graph_trace_ef$V.txt
done
