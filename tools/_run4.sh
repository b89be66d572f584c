export ACE_B200_DATA_FILE=$(python -c "import bench; print(bench.weight_file('resnet20_cifar10_pre'))") RTLIB_BTS_EVEN_POLY=1 ACE_B200_QUIET=1
ACE_B200_STATS=1 tests/_emitted_bin/resnet20_cifar10_pre 6 2>&1 | grep -E "driver\] (image|Prep)|stats\] (kernels|Boot|Alloc)"
timeout 900 python bench.py --no-cpu > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; cut -c1-400 gpurun_out/bench_d.json
