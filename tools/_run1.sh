set -x
ACE_MODEL_PARITY=0 timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_a.log
cat gpurun_out/pytest_a.log
tools/gpu_profile_run.sh a
timeout 900 python bench.py > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; cat gpurun_out/bench_a.json | cut -c1-600
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ntt_|base_conv|ksw_inner' -s 16 -c 10 -f -o gpurun_out/full_a python tools/microbench.py 34 > gpurun_out/ncu_a.log 2>&1
tail -3 gpurun_out/ncu_a.log
