set -x
export ACE_MODEL_PARITY=0
timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_client.py -x -q 2>&1 | tail -5
tools/gpu_profile_run.sh c
grep -E "driver|stats\]" gpurun_out/stats_c.log | grep -v logits
