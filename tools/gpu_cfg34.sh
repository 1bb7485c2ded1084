#!/bin/bash
# BASELINE.json configs 3 (D-D-only net) and 4 (scaled graph) on N GPUs; usage: bash tools/gpu_cfg34.sh <tag> <N>
TAG=$1; N=$2
O=gpurun_out; mkdir -p $O
run() {  # name, extra args
  NAME=$1; shift
  PORT=$((29500 + RANDOM % 1000))
  if [ "$N" = "1" ]; then
    timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 3 --skip-cpu-baseline "$@" > $O/${TAG}_${NAME}_n${N}.json 2> $O/${TAG}_${NAME}_n${N}.err
  else
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 20 --warmup 3 "$@" > $O/${TAG}_${NAME}_n${N}.json 2> $O/${TAG}_${NAME}_n${N}.err
  fi
  echo "$NAME N=$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_${NAME}_n${N}.json").read().strip().splitlines()[-1])
    print("  ms_per_step %.3f value %.3g e2e_ms %s loss %s edges %s" % (d["ms_per_step"], d["value"], d["e2e"] and round(d["e2e"]["ms_per_step"],2), d.get("loss"), d["config"]["directed_dd_edges"]))
except Exception as e:
    print("  no json line:", e)
PY
  grep -v "Warning\|warn" $O/${TAG}_${NAME}_n${N}.err | tail -4
}
run dd --model dd
if [ "$3" = "scaled" ]; then run scaled --shape scaled; fi
