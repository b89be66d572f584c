export ACE_MODEL_PARITY=1
timeout 3000 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_e.log; cat gpurun_out/pytest_e.log
timeout 900 python bench.py > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; cut -c1-300 gpurun_out/bench_e.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_e_ref.json 2>> gpurun_out/bench_e.err; cut -c1-300 gpurun_out/bench_e_ref.json
