#!/bin/bash
# One GPU session: model runs with device-time accounting (ACE_B200_PROF) and host-time stats,
# bench line, ncu full capture of the transform kernels.  Output under gpurun_out/.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-a}
MSG=$(python -c "import bench; print(bench.weight_file('resnet20_cifar10_pre'))")
export ACE_B200_DATA_FILE=$MSG RTLIB_BTS_EVEN_POLY=1 ACE_B200_QUIET=1
BIN=tests/_emitted_bin/resnet20_cifar10_pre
ACE_B200_PROF=1 timeout 600 $BIN 2 > gpurun_out/prof_$TAG.log 2>&1
ACE_B200_STATS=1 timeout 600 $BIN 2 > gpurun_out/stats_$TAG.log 2>&1
tail -3 gpurun_out/stats_$TAG.log
