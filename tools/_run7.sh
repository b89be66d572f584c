for pin in 0 30 60; do
  echo "== L2PIN=$pin"
  ACE_B200_L2PIN=$pin python tools/microbench.py 34 17 2>&1 | grep -E "L2 window|ntt|intt|key_switch|ct_rotate|mul_relin|mod_down|modup"
done
export ACE_B200_DATA_FILE=$(python -c "import bench; print(bench.weight_file('resnet20_cifar10_pre'))") RTLIB_BTS_EVEN_POLY=1 ACE_B200_QUIET=1
for pin in 0 40; do
  echo "== model L2PIN=$pin"
  ACE_B200_L2PIN=$pin tests/_emitted_bin/resnet20_cifar10_pre 5 2>&1 | grep -E "L2 window|driver\] image"
done
