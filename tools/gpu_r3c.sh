#!/bin/bash
# persistent tensor-core node kernels: parity, A/B timing against the per-node kernels, one full ncu capture with source
TAG=${1:-rXX}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "rgcn or benched or tip_model or dd_net" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log
for V in 2 0; do
TIPB_RGCN_TCP=$V timeout 200 python tools/ubench_rgcn.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('TCP=$V', {k: v for k, v in d['us'].items() if 'node' in k})"
done
python -c "from tip_b200 import _lib; print('tc status', _lib.lib().tipb_rgcn_tc_status())"
bash tools/gpu_ncu_cmd.sh ${TAG} "k_rgcn_node_(fwd|bwd)_tcp" 4 2 python tools/ubench_rgcn.py
