#!/bin/bash
# compute-sanitizer over the library kernels: memcheck on a fast subset of the GPU parity tests and on smoke(), racecheck
# (shared-memory hazards: the per-warp smem slots of k_seg_aggregate_flat, the tile sort of k_pair_pass, the look-back
# scan) on smoke() and the small-shape operator tests.  usage (under gpurun): bash tools/gpu_sanitizer.sh <tag>
TAG=${1:-rXX}
O=gpurun_out; mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
SMALL="rgcn_conv2_matches or hierarchy_conv or pp_encoder or decoder_matches or typed_csr_bit_exact or bce_loss or nn_decoder or hier_encoder or pair_pass_against or process_edges_on_the_device_matches or negative_sampling_bit_exact"
timeout 1500 $CS --tool memcheck --error-exitcode 9 --log-file $O/${TAG}_memcheck.log \
    python -m pytest tests -m gpu -q -x -k "$SMALL" > $O/${TAG}_memcheck_pytest.log 2>&1; echo "memcheck rc=$?"
tail -3 $O/${TAG}_memcheck_pytest.log; grep -c "Invalid\|out of bounds" $O/${TAG}_memcheck.log; tail -3 $O/${TAG}_memcheck.log
timeout 900 $CS --tool racecheck --racecheck-report all --error-exitcode 9 --log-file $O/${TAG}_racecheck.log \
    python __graft_entry__.py smoke > $O/${TAG}_racecheck_smoke.log 2>&1; echo "racecheck(smoke) rc=$?"
tail -2 $O/${TAG}_racecheck_smoke.log; grep -c "hazard" $O/${TAG}_racecheck.log; tail -3 $O/${TAG}_racecheck.log
timeout 1500 $CS --tool racecheck --racecheck-report all --error-exitcode 9 --log-file $O/${TAG}_racecheck_ops.log \
    python -m pytest tests -m gpu -q -x -k "rgcn_conv2_matches or pair_pass_against or bce_loss_against or typed_csr_bit_exact" > $O/${TAG}_racecheck_pytest.log 2>&1; echo "racecheck(ops) rc=$?"
tail -3 $O/${TAG}_racecheck_pytest.log; grep -c "hazard" $O/${TAG}_racecheck_ops.log; tail -3 $O/${TAG}_racecheck_ops.log
timeout 900 $CS --tool initcheck --error-exitcode 9 --log-file $O/${TAG}_initcheck.log \
    python __graft_entry__.py smoke > $O/${TAG}_initcheck_smoke.log 2>&1; echo "initcheck(smoke) rc=$?"
grep -c "Uninitialized" $O/${TAG}_initcheck.log; tail -3 $O/${TAG}_initcheck.log
