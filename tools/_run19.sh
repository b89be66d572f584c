for S in 4 6; do
timeout 900 python bench.py --no-cpu --steps 3 --streams $S > gpurun_out/bench_s$S.json 2> gpurun_out/bench_s$S.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_s$S.json'))
    print("streams $S", d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['s_per_image'], d['gpu_launches'])
except Exception as e:
    print("streams $S failed", e); print(open('gpurun_out/bench_s$S.err').read()[-1500:])
PY
nvidia-smi --query-gpu=memory.used --format=csv,noheader
done
