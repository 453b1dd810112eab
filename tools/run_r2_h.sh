# round 2, GPU call H (1 GPU): full GPU test suite, default bench (as the driver runs it), reference arms, ncu of the shipped kernels,
# compute-sanitizer passes, smoke
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2h_pytest.log 2>&1; tail -12 gpurun_out/r2h_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2h_smoke.log 2>&1; tail -2 gpurun_out/r2h_smoke.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; tail -2 gpurun_out/r2h_bench_n1.err; cut -c1-400 gpurun_out/r2h_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2h_bench_ref.json 2> gpurun_out/r2h_bench_ref.err; cut -c1-300 gpurun_out/r2h_bench_ref.json
timeout 600 python bench.py --impl reference-gpu --steps 3 --warmup 2 > gpurun_out/r2h_bench_refgpu_1024.json 2> gpurun_out/r2h_bench_refgpu_1024.err; cut -c1-200 gpurun_out/r2h_bench_refgpu_1024.json
timeout 900 python bench.py --impl reference-gpu --resolution 3840 --steps 1 --warmup 1 > gpurun_out/r2h_bench_refgpu_3840.json 2> gpurun_out/r2h_bench_refgpu_3840.err; tail -2 gpurun_out/r2h_bench_refgpu_3840.err; cut -c1-200 gpurun_out/r2h_bench_refgpu_3840.json
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/r2h_kernels python tools/ncu_kernels.py attn,attn3840,gn,geglu,ln,publish,linear > gpurun_out/r2h_ncu.log 2>&1; tail -2 gpurun_out/r2h_ncu.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 2400 --csv --log-file gpurun_out/r2h_launches_bench_window.csv python bench.py --steps 1 --warmup 3 --no-cuda-graph --no-cpu-baseline --no-hires --no-roofline > gpurun_out/r2h_ncu_bench.log 2>&1; tail -1 gpurun_out/r2h_ncu_bench.log | cut -c1-200
timeout 300 compute-sanitizer --tool racecheck --print-limit 3 python -m pytest tests/test_kernels_gpu.py -q -x -k "single_segment and (256-384 or 200-77-2-64) or groupnorm_local and 320 or fused_halo and True" > gpurun_out/r2h_sanitizer_racecheck.log 2>&1; tail -6 gpurun_out/r2h_sanitizer_racecheck.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_linear_gpu.py tests/test_kernels_gpu.py -q -x -k "epilogues and 256-256-64 or geglu_fused and 200 or publish_and_wait or halo_push" > gpurun_out/r2h_sanitizer_memcheck.log 2>&1; tail -6 gpurun_out/r2h_sanitizer_memcheck.log
