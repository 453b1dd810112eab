# round 2, GPU call C: attention v4 (software-pipelined) tests + variant sweep with graph timing, GroupNorm fused (+halo) tests,
# GEMM graph timing + ncu, PDL on/off
mkdir -p gpurun_out
DF_PDL=0 timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_linear_gpu.py -q > gpurun_out/r2c_kernel_tests_nopdl.log 2>&1; tail -6 gpurun_out/r2c_kernel_tests_nopdl.log
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_linear_gpu.py -q > gpurun_out/r2c_kernel_tests.log 2>&1; tail -6 gpurun_out/r2c_kernel_tests.log
timeout 900 python -m pytest tests/test_unet_gpu.py -q -k "not full_size" > gpurun_out/r2c_unet_tests.log 2>&1; tail -8 gpurun_out/r2c_unet_tests.log
timeout 300 python tools/bench_linear.py 1024 > gpurun_out/r2c_linear_vs_cublas.txt 2>&1; cat gpurun_out/r2c_linear_vs_cublas.txt
SH="1024_l1,1024_l2,3840n4_l2,1024n4_l2,1024n4_l1,cross_l2"
rm -f gpurun_out/r2c_attn_sweep.txt
for V in "-DDF_FMHA_PIPE=1 -DDF_PIPE_PREFETCH_AT=2" "-DDF_FMHA_PIPE=1 -DDF_PIPE_PREFETCH_AT=0" "-DDF_FMHA_PIPE=1 -DDF_PIPE_PREFETCH_AT=2 -DDF_EMU_QUARTERS=0" "-DDF_FMHA_PIPE=1 -DDF_PIPE_PREFETCH_AT=2 -DDF_EMU_QUARTERS=2" "-DDF_FMHA_PIPE=0"; do
  echo "== variant: $V" >> gpurun_out/r2c_attn_sweep.txt
  DF_NVCC_FLAGS="$V" python -c "from distrifuser_b200 import build; build.build(force=True)" >> gpurun_out/r2c_build.log 2>&1
  timeout 300 python tools/bench_attn.py --shapes $SH >> gpurun_out/r2c_attn_sweep.txt 2>&1
done
python -c "from distrifuser_b200 import build; build.build(force=True)" >> gpurun_out/r2c_build.log 2>&1
cat gpurun_out/r2c_attn_sweep.txt
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/r2c_kernels python tools/ncu_kernels.py attn,attn3840,gn,linear > gpurun_out/r2c_ncu.log 2>&1; tail -3 gpurun_out/r2c_ncu.log
python tools/bench_vs_torch.py > gpurun_out/r2c_vs_torch.txt 2>&1; cat gpurun_out/r2c_vs_torch.txt
for P in 1 0; do DF_PDL=$P timeout 600 python bench.py --steps 3 --warmup 3 --no-hires --no-cpu-baseline > gpurun_out/r2c_bench_pdl$P.json 2> gpurun_out/r2c_bench_pdl$P.err; tail -2 gpurun_out/r2c_bench_pdl$P.err; python -c "
import json; d=json.load(open('gpurun_out/r2c_bench_pdl$P.json')); print('PDL=$P', d['value'], d['roofline']['achieved'], [ (s['shape']['lq'], round(s['tflops'])) for s in d['roofline']['shapes']], d['roofline']['groupnorm']['achieved'])"; done
