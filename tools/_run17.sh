timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
python tools/microbench.py 34 17 2>&1 | grep -E "ntt|intt|modup|mod_down|key_switch|ct_rotate|mul_relin"
export ACE_B200_DATA_FILE=$(python -c "import bench; print(bench.weight_file('resnet20_cifar10_pre'))") RTLIB_BTS_EVEN_POLY=1 ACE_B200_QUIET=1
tests/_emitted_bin/resnet20_cifar10_pre 4 2>&1 | grep -E "driver\] (image|logits 3)"
