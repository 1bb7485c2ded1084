#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, ncu --set full of the hot kernels.
# usage (from repo root, under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?"
cat $OUT/${TAG}_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_eager.csv \
    python tools/one_step.py 3 > $OUT/${TAG}_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k 'regex:k_seg_aggregate|k_decoder_seg|k_node_aggregate|k_grp_scatter|k_mt_generate|k_rgcn_node|k_chain_walk|k_window_scan' \
    -s 26 -c 24 -f -o $OUT/${TAG}_full python tools/one_step.py 2 > $OUT/${TAG}_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
