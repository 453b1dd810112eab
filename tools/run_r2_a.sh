# round 2, GPU call A: full GPU test suite (new real-size parity cases), bench (ours / reference / reference-gpu), ncu of shipped kernels
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r2a_pytest.log 2>&1; tail -25 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; tail -3 gpurun_out/r2a_bench_n1.err; cat gpurun_out/r2a_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 --cpu-budget-s 150 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; tail -3 gpurun_out/r2a_bench_ref.err; cat gpurun_out/r2a_bench_ref.json
timeout 900 python bench.py --impl reference-gpu --steps 3 --warmup 2 > gpurun_out/r2a_bench_refgpu.json 2> gpurun_out/r2a_bench_refgpu.err; tail -5 gpurun_out/r2a_bench_refgpu.err; cat gpurun_out/r2a_bench_refgpu.json
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/r2a_kernels python tools/ncu_kernels.py attn,attn3840,gn,geglu,ln,publish > gpurun_out/r2a_ncu.log 2>&1; tail -3 gpurun_out/r2a_ncu.log
python tools/bench_vs_torch.py > gpurun_out/r2a_vs_torch.txt 2>&1; cat gpurun_out/r2a_vs_torch.txt
ls -la gpurun_out/*.ncu-rep
