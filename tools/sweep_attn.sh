mkdir -p gpurun_out; rm -f gpurun_out/attn_sweep.txt
timeout 120 python tools/bench_attn.py --shapes 1024_l2 --kernel 2 --iters 2 > gpurun_out/fmha2_first.txt 2>&1; tail -3 gpurun_out/fmha2_first.txt
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" > gpurun_out/k14.log 2>&1; tail -6 gpurun_out/k14.log
for K in 1 2; do echo "== kernel $K" >> gpurun_out/attn_sweep.txt; timeout 120 python tools/bench_attn.py --shapes 1024_l1,1024_l2,3840n4_l2,2048n2_l1 --kernel $K >> gpurun_out/attn_sweep.txt 2>&1; done
cat gpurun_out/attn_sweep.txt
