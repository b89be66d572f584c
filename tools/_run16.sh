export ACE_MODEL_PARITY=1
timeout 3400 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_g.log; cat gpurun_out/pytest_g.log
timeout 900 python bench.py > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; cut -c1-260 gpurun_out/bench_g.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_g_ref.json 2>> gpurun_out/bench_g.err; cut -c1-200 gpurun_out/bench_g_ref.json
