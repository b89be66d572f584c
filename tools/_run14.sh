export ACE_B200_DATA_FILE=$(python -c "import bench; print(bench.weight_file('resnet20_cifar10_pre'))") RTLIB_BTS_EVEN_POLY=1 ACE_B200_QUIET=1
BIN=tests/_emitted_bin/resnet20_cifar10_pre
# launch list: a window of 12000 launches inside the first image (after key generation)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 12000 --csv --log-file gpurun_out/launches_v3.csv $BIN 1 > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
# full capture of the top kernels inside the run
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'base_conv|ew_chain|ksw_inner_rot|ntt_inv_tile8|pt_dot' -s 3000 -c 10 -f -o gpurun_out/full_v3 $BIN 1 > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
gzip -f gpurun_out/launches_v3.csv
