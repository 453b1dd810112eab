# round 2, GPU call B: TMEM fragment probe, tcgen05 GEMM tests + vs-cuBLAS table, attention v3 tests + variant sweep
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tmem_probe tools/probes/tmem_layout_probe.cu && timeout 60 gpurun_out/tmem_probe > gpurun_out/r2b_tmem_probe.txt 2>&1; head -12 gpurun_out/r2b_tmem_probe.txt
timeout 600 python -m pytest tests/test_linear_gpu.py -q -x > gpurun_out/r2b_linear_tests.log 2>&1; tail -15 gpurun_out/r2b_linear_tests.log
timeout 300 python tools/bench_linear.py 1024 > gpurun_out/r2b_linear_vs_cublas.txt 2>&1; cat gpurun_out/r2b_linear_vs_cublas.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k attention > gpurun_out/r2b_attn_tests.log 2>&1; tail -8 gpurun_out/r2b_attn_tests.log
SH="1024_l1,1024_l2,3840n4_l2,1024n4_l2,1024n4_l1"
for V in "-DDF_FMHA_V3=1 -DDF_EMU_QUARTERS=1" "-DDF_FMHA_V3=1 -DDF_EMU_QUARTERS=0" "-DDF_FMHA_V3=1 -DDF_EMU_QUARTERS=2" "-DDF_FMHA_V3=0"; do
  echo "== variant: $V" >> gpurun_out/r2b_attn_sweep.txt
  DF_NVCC_FLAGS="$V" python -c "from distrifuser_b200 import build; build.build(force=True)" >> gpurun_out/r2b_build.log 2>&1
  timeout 300 python tools/bench_attn.py --shapes $SH >> gpurun_out/r2b_attn_sweep.txt 2>&1
done
python -c "from distrifuser_b200 import build; build.build(force=True)" >> gpurun_out/r2b_build.log 2>&1
cat gpurun_out/r2b_attn_sweep.txt
