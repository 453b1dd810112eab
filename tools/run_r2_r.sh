# round 2, GPU call R (1 GPU): final build -- full GPU suite, smoke, the default bench as the driver runs it, ncu captures of the
# shipped kernels and the launch list of the bench command, SASS-free
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -10 gpurun_out/r2r_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2r_smoke.log 2>&1; tail -1 gpurun_out/r2r_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2r_bench_n1.json 2> gpurun_out/r2r_bench_n1.err; tail -2 gpurun_out/r2r_bench_n1.err; cut -c1-260 gpurun_out/r2r_bench_n1.json
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/r2r_kernels python tools/ncu_kernels.py attn,attn3840,gn,geglu,ln,bias,publish,linear > gpurun_out/r2r_ncu.log 2>&1; tail -2 gpurun_out/r2r_ncu.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 2200 --csv --log-file gpurun_out/r2r_launches_bench_window.csv python bench.py --steps 1 --warmup 3 --no-cuda-graph --no-cpu-baseline --no-hires --no-roofline > gpurun_out/r2r_ncu_bench.log 2>&1; tail -1 gpurun_out/r2r_ncu_bench.log | cut -c1-200
