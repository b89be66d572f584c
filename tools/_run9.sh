export RTLIB_BTS_EVEN_POLY=1 ACE_B200_QUIET=1 ACE_B200_DEBUG_BTS=1
M=resnet110_cifar10_train
for amp in 0.02 0.005; do
echo "== amp $amp"
ACE_B200_DATA_FILE=$(python -c "import bench; print(bench.weight_file('$M', $amp))") tests/_emitted_bin/$M 1 2>&1 | grep -E "bts|driver" > gpurun_out/r110_dbg_$amp.log
head -9 gpurun_out/r110_dbg_$amp.log | tail -7; tail -3 gpurun_out/r110_dbg_$amp.log
done
