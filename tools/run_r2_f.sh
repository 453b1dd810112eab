# round 2, GPU call F: balanced-tail attention schedule (tests + graph-timed shapes), GEMM GEGLU 256-wide + 8 epilogue warps, N=1 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_linear_gpu.py -q > gpurun_out/r2f_kernel_tests.log 2>&1; tail -8 gpurun_out/r2f_kernel_tests.log
timeout 300 python tools/bench_attn.py --shapes 1024_l1,1024_l2,3840n4_l2,1024n4_l2,1024n4_l1,1024n2_l2,1024n2_l1,cross_l2,sd15_l0,sd15_l1 > gpurun_out/r2f_attn.txt 2>&1; cat gpurun_out/r2f_attn.txt
DF_NVCC_FLAGS="-DDF_MIN_PART_TILES=4" python -c "from distrifuser_b200 import build; build.build(force=True)" >> gpurun_out/r2f_build.log 2>&1
echo "== DF_MIN_PART_TILES=4" >> gpurun_out/r2f_attn.txt; timeout 300 python tools/bench_attn.py --shapes 1024_l1,1024_l2,1024n4_l2,1024n4_l1,1024n2_l2 >> gpurun_out/r2f_attn.txt 2>&1
DF_NVCC_FLAGS="-DDF_MIN_PART_TILES=1" python -c "from distrifuser_b200 import build; build.build(force=True)" >> gpurun_out/r2f_build.log 2>&1
echo "== DF_MIN_PART_TILES=1" >> gpurun_out/r2f_attn.txt; timeout 300 python tools/bench_attn.py --shapes 1024_l1,1024_l2,1024n4_l2,1024n4_l1,1024n2_l2 >> gpurun_out/r2f_attn.txt 2>&1
python -c "from distrifuser_b200 import build; build.build(force=True)" >> gpurun_out/r2f_build.log 2>&1
tail -14 gpurun_out/r2f_attn.txt
timeout 300 python tools/bench_linear.py 1024 > gpurun_out/r2f_linear_vs_cublas.txt 2>&1; tail -4 gpurun_out/r2f_linear_vs_cublas.txt
python tools/bench_vs_torch.py > gpurun_out/r2f_vs_torch.txt 2>&1; head -6 gpurun_out/r2f_vs_torch.txt
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_pipeline_gpu.py -q -k "not full_size" > gpurun_out/r2f_unet_tests.log 2>&1; tail -4 gpurun_out/r2f_unet_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-hires --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -2 gpurun_out/r2f_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2f_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], [ (s['shape']['lq'], round(s['tflops'])) for s in d['roofline']['shapes']], d['roofline']['groupnorm']['achieved'])"
