#!/bin/bash
# compute-sanitizer over the tensor-core kernels (R-GCN node kernels at >= 256 relations, decoder sweep) and the pair pass
TAG=${1:-rXX}
O=gpurun_out; mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
K="rgcn_against_fp64_oracle or decoder_sweep_all_widths or second_gcn"
timeout 280 $CS --tool memcheck --error-exitcode 9 --log-file $O/${TAG}_memcheck_tc.log python -m pytest tests -m gpu -q -x -k "$K" > $O/${TAG}_memcheck_tc.out 2>&1; echo "memcheck(tc) rc=$?"
tail -2 $O/${TAG}_memcheck_tc.out; tail -2 $O/${TAG}_memcheck_tc.log
timeout 280 $CS --tool racecheck --racecheck-report all --error-exitcode 9 --log-file $O/${TAG}_racecheck_ops.log python -m pytest tests -m gpu -q -x -k "rgcn_against_fp64_oracle or pair_pass_against or bce_loss_against or negative_sampling_bit_exact" > $O/${TAG}_racecheck_ops.out 2>&1; echo "racecheck(ops) rc=$?"
tail -2 $O/${TAG}_racecheck_ops.out; grep -c "hazard" $O/${TAG}_racecheck_ops.log; tail -3 $O/${TAG}_racecheck_ops.log
timeout 200 $CS --tool initcheck --error-exitcode 9 --log-file $O/${TAG}_initcheck_smoke.log python __graft_entry__.py smoke > $O/${TAG}_initcheck_smoke.out 2>&1; echo "initcheck(smoke) rc=$?"
tail -1 $O/${TAG}_initcheck_smoke.out; tail -2 $O/${TAG}_initcheck_smoke.log
