ACE_MODEL_PARITY=0 timeout 900 python -m pytest tests/test_gpu_model.py -x -q -k "resnet20" 2>&1 | tail -3
for S in 1 2 3; do
timeout 900 python bench.py --no-cpu --steps 3 --streams $S > gpurun_out/bench_s$S.json 2> gpurun_out/bench_s$S.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_s$S.json'))
    print("streams $S", d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['s_per_image'], d['gpu_launches'], d['config']['logits0'])
except Exception as e:
    print("streams $S failed", e); print(open('gpurun_out/bench_s$S.err').read()[-1500:])
PY
done
