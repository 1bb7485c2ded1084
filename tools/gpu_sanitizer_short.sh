#!/bin/bash
# compute-sanitizer, time-boxed: memcheck over smoke() and the small-shape operator tests, racecheck over smoke()
# (shared-memory hazards: per-warp slots of k_seg_aggregate_flat, tile sort of k_pair_pass, window scan masks, look-back scan)
TAG=${1:-rXX}
O=gpurun_out; mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 200 $CS --tool memcheck --error-exitcode 9 --log-file $O/${TAG}_memcheck_smoke.log python __graft_entry__.py smoke > $O/${TAG}_memcheck_smoke.out 2>&1; echo "memcheck(smoke) rc=$?"
tail -1 $O/${TAG}_memcheck_smoke.out; tail -2 $O/${TAG}_memcheck_smoke.log
timeout 200 $CS --tool racecheck --racecheck-report all --error-exitcode 9 --log-file $O/${TAG}_racecheck_smoke.log python __graft_entry__.py smoke > $O/${TAG}_racecheck_smoke.out 2>&1; echo "racecheck(smoke) rc=$?"
tail -1 $O/${TAG}_racecheck_smoke.out; grep -c "hazard" $O/${TAG}_racecheck_smoke.log; tail -2 $O/${TAG}_racecheck_smoke.log
SMALL="rgcn_conv2_matches or hierarchy_conv or pp_encoder or decoder_matches or bce_loss_against or pair_pass_against or negative_sampling_bit_exact"
timeout 280 $CS --tool memcheck --error-exitcode 9 --log-file $O/${TAG}_memcheck_ops.log python -m pytest tests -m gpu -q -x -k "$SMALL" > $O/${TAG}_memcheck_ops.out 2>&1; echo "memcheck(ops) rc=$?"
tail -2 $O/${TAG}_memcheck_ops.out; tail -2 $O/${TAG}_memcheck_ops.log
