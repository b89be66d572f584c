export RTLIB_BTS_EVEN_POLY=1 ACE_B200_QUIET=1 ACE_B200_DEBUG_BTS=1 ACE_B200_DEBUG_RANGE=100000
for M in resnet110_cifar10_train resnet56_cifar10_pre; do
ACE_B200_DATA_FILE=$(python -c "import bench; print(bench.weight_file('$M'))") timeout 600 tests/_emitted_bin/$M 1 2>&1 | grep -E "bts|range" | uniq -c | awk '/bts   6/{exit} {print}' > gpurun_out/range_$M.log
wc -l gpurun_out/range_$M.log
done
