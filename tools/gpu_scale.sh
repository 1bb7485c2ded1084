#!/bin/bash
# one N-GPU visit: config 1 (TIP-cat) bench line + rank-0 graph trace, config 3 (D-D-only net), config 4 (scaled graph)
# usage (under gpurun --gpus N): bash tools/gpu_scale.sh <tag> <N> [noscaled]
TAG=$1; N=$2
O=gpurun_out; mkdir -p $O
launch() { PORT=$((29500 + RANDOM % 1000)); python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT "$@"; }
summary() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  N=%d ms_per_step %.3f value %.4g e2e_ms %s loss %s edges %s" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"] and round(d["e2e"]["ms_per_step"], 2), d.get("loss"), d["config"]["directed_dd_edges"]))
    print("  families", [(f["family"][:14], f["us"]) for f in d["roofline"]["families"]][:8])
except Exception as e:
    print("  no json line:", e)
PY
}
timeout 900 bash -c "$(declare -f launch); N=$N; launch bench.py --gpus $N --steps 30 --warmup 5" > $O/${TAG}_bench_n${N}.json 2> $O/${TAG}_bench_n${N}.err; echo "config 1 N=$N rc=$?"; summary $O/${TAG}_bench_n${N}.json
timeout 600 bash -c "$(declare -f launch); N=$N; launch tools/graph_trace_multi.py $O/${TAG}_trace_n${N}.txt" > $O/${TAG}_trace_n${N}.log 2>&1; echo "trace rc=$?"; head -1 $O/${TAG}_trace_n${N}.txt
timeout 900 bash -c "$(declare -f launch); N=$N; launch bench.py --gpus $N --steps 20 --warmup 3 --model dd" > $O/${TAG}_dd_n${N}.json 2> $O/${TAG}_dd_n${N}.err; echo "config 3 (dd) N=$N rc=$?"; summary $O/${TAG}_dd_n${N}.json
if [ "$3" != "noscaled" ]; then
timeout 1500 bash -c "$(declare -f launch); N=$N; launch bench.py --gpus $N --steps 10 --warmup 3 --shape scaled" > $O/${TAG}_scaled_n${N}.json 2> $O/${TAG}_scaled_n${N}.err; echo "config 4 (scaled) N=$N rc=$?"; summary $O/${TAG}_scaled_n${N}.json
grep -v "Warning\|warn" $O/${TAG}_scaled_n${N}.err | tail -4
fi
