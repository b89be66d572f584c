export RTLIB_BTS_EVEN_POLY=1 ACE_B200_QUIET=1 ACE_B200_DEBUG_BTS=1
M=resnet110_cifar10_train
ACE_B200_DATA_FILE=$(python -c "import bench; print(bench.weight_file('$M'))") tests/_emitted_bin/$M 1 2>&1 | grep -E "bts|driver" | head -130 > gpurun_out/r110_dbg.log
head -30 gpurun_out/r110_dbg.log; tail -4 gpurun_out/r110_dbg.log
M=resnet56_cifar10_pre
ACE_B200_DATA_FILE=$(python -c "import bench; print(bench.weight_file('$M'))") tests/_emitted_bin/$M 1 2>&1 | grep -E "bts|driver" > gpurun_out/r56_dbg.log
head -8 gpurun_out/r56_dbg.log; tail -4 gpurun_out/r56_dbg.log
