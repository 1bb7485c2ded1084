#!/bin/bash
# all GPU parity tests; usage: bash tools/gpu_pytest.sh <tag> [pytest args]
TAG=${1:-rXX}; shift
O=gpurun_out; mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -q --durations=8 "$@" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
grep -n "^E  \|passed\|failed\|^FAILED\|^ERROR\|s call" $O/${TAG}_pytest.log | head -40
