# round 2, GPU call V (1 GPU): last full GPU suite + smoke + short bench on the final tree
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=3 > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -7 gpurun_out/r2v_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2v_smoke.log 2>&1; tail -1 gpurun_out/r2v_smoke.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-hires --no-cpu-baseline --no-roofline > gpurun_out/r2v_bench_n1.json 2> gpurun_out/r2v_bench_n1.err; cut -c1-200 gpurun_out/r2v_bench_n1.json
