O=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__block_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,lts__t_sector_hit_rate.pct \
  --clock-control none -k 'regex:^k_' -s 215 -c 120 --csv --page raw --log-file $O/r01l_allkernels_raw.csv python tools/one_step.py 3 > $O/r01l_all.log 2>&1; echo "metrics rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_seg_aggregate_flat|k_decoder_seg|k_grp_place' -s 18 -c 7 -f -o $O/r01l_top python tools/one_step.py 3 > $O/r01l_top.log 2>&1; echo "full rc=$?"
ncu -i $O/r01l_top.ncu-rep --page raw --csv > $O/r01l_top_raw.csv 2>/dev/null
for k in k_seg_aggregate_flat k_decoder_seg k_grp_place; do ncu -i $O/r01l_top.ncu-rep --page source --csv --kernel-name regex:$k --launch-count 1 > $O/r01l_src_$k.csv 2>/dev/null; done
ls -la $O; du -sm $O
