timeout 1500 python -m pytest tests/test_gpu_bootstrap.py -x -q 2>&1 | tail -25 > gpurun_out/pytest_f.log; cat gpurun_out/pytest_f.log
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_client.py -x -q 2>&1 | tail -5
tools/gpu_profile_run.sh f
