#!/bin/bash
# multi-GPU bench lines; usage: bash tools/gpu_multi.sh <tag> <N> [extra bench args]
TAG=$1; N=$2; shift; shift
O=gpurun_out; mkdir -p $O
PORT=$((29500 + RANDOM % 1000))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 30 --warmup 5 "$@" > $O/${TAG}_bench_n${N}.json 2> $O/${TAG}_bench_n${N}.err; echo "bench N=$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_bench_n${N}.json").read().strip().splitlines()[-1])
    print("N=%d ms_per_step %.3f value %.3g e2e_ms %s loss %s" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"] and round(d["e2e"]["ms_per_step"],2), d.get("loss")))
    print("dominant", d["roofline"]["kernel"], d["roofline"]["kernel_us"], [ (f["family"][:12], f["us"]) for f in d["roofline"]["families"]])
except Exception as e:
    print("no json line:", e)
PY
tail -5 $O/${TAG}_bench_n${N}.err
