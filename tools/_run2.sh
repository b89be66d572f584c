set -x
export ACE_MODEL_PARITY=0
timeout 900 python -m pytest tests/test_gpu_client.py -x -q 2>&1 | tail -5
timeout 1200 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -15
tools/gpu_profile_run.sh b
grep -E "driver|stats\]" gpurun_out/stats_b.log | grep -v logits
