ACE_MODEL_PARITY=1 timeout 2500 python -m pytest tests/test_gpu_model.py -x -q -k "bit_exact" 2>&1 | tail -12
