#!/bin/bash
# round-2 visit A: all GPU parity tests (incl. the benched workload), a bench line, the full CPU reference step on the box
TAG=${1:-r02a}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/${TAG}_smi.txt; nproc >> $O/${TAG}_smi.txt; free -g >> $O/${TAG}_smi.txt
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^E  \|passed\|failed\|^FAILED\|s call" $O/${TAG}_pytest.log | head -30
timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench rc=$?"
cut -c1-400 $O/${TAG}_bench_n1.json
timeout 900 python tools/cpu_full_step.py $O/${TAG}_cpu_full_step_box.json > $O/${TAG}_cpu_full.log 2>&1; echo "cpu full rc=$?"; cat $O/${TAG}_cpu_full_step_box.json
