ACE_MODEL_PARITY=0 timeout 900 python -m pytest tests/test_gpu_model.py -x -q -k "three_images" 2>&1 | tail -5
timeout 900 python bench.py --no-cpu --steps 3 --streams 4 > gpurun_out/bench_s4.json 2> gpurun_out/bench_s4.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_s4.json'))
    print("streams 4", d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['s_per_image'], d['gpu_launches'])
except Exception as e:
    print("streams 4 failed", e); print(open('gpurun_out/bench_s4.err').read()[-800:])
PY
