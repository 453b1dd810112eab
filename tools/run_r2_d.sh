# round 2, GPU call D: GEMM with 8 epilogue warps / GEGLU 160-wide tiles, attention v3 vs v4', bench with graph-timed GroupNorm
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_linear_gpu.py -q > gpurun_out/r2d_kernel_tests.log 2>&1; tail -6 gpurun_out/r2d_kernel_tests.log
timeout 300 python tools/bench_linear.py 1024 > gpurun_out/r2d_linear_vs_cublas.txt 2>&1; cat gpurun_out/r2d_linear_vs_cublas.txt
SH="1024_l1,1024_l2,3840n4_l2,1024n4_l2,1024n4_l1"
rm -f gpurun_out/r2d_attn_sweep.txt
for V in "-DDF_FMHA_PIPE=1" "-DDF_FMHA_PIPE=1 -DDF_EMU_QUARTERS=0" "-DDF_FMHA_PIPE=0"; do
  echo "== variant: $V" >> gpurun_out/r2d_attn_sweep.txt
  DF_NVCC_FLAGS="$V" python -c "from distrifuser_b200 import build; build.build(force=True)" >> gpurun_out/r2d_build.log 2>&1
  timeout 300 python tools/bench_attn.py --shapes $SH >> gpurun_out/r2d_attn_sweep.txt 2>&1
done
python -c "from distrifuser_b200 import build; build.build(force=True)" >> gpurun_out/r2d_build.log 2>&1
cat gpurun_out/r2d_attn_sweep.txt
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/r2d_kernels python tools/ncu_kernels.py attn,linear > gpurun_out/r2d_ncu.log 2>&1; tail -3 gpurun_out/r2d_ncu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-hires --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -2 gpurun_out/r2d_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2d_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], [ (s['shape']['lq'], round(s['tflops'])) for s in d['roofline']['shapes']], d['roofline']['groupnorm'])"
DF_LINEAR=none timeout 600 python bench.py --steps 3 --warmup 3 --no-hires --no-cpu-baseline --no-roofline > gpurun_out/r2d_bench_nolinear.json 2> gpurun_out/r2d_bench_nolinear.err; python -c "
import json; d=json.load(open('gpurun_out/r2d_bench_nolinear.json')); print('DF_LINEAR=none', d['value'], d['e2e']['value'])"
