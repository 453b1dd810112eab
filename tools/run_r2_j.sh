# round 2, GPU call J (1 GPU): speculative-reference softmax -- parity tests, then graph-timed A/B against the previous kernel
# (variants prebuilt in distrifuser_b200/variants by build.build(out=...))
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" > gpurun_out/r2j_attn_tests.log 2>&1; tail -5 gpurun_out/r2j_attn_tests.log
SH="1024_l1,1024_l2,3840n4_l2,3840n4_l1,1024n4_l1,sd15_l0"
rm -f gpurun_out/r2j_attn_sweep.txt
for V in ctl emu4 emu4_opq0 emu3 emu5 emu6 emu6_opq0 emu8; do
  echo "== variant: $V" >> gpurun_out/r2j_attn_sweep.txt
  DF_LIB_PATH=distrifuser_b200/variants/lib_$V.so timeout 200 python tools/bench_attn.py --shapes $SH >> gpurun_out/r2j_attn_sweep.txt 2>&1
done
cat gpurun_out/r2j_attn_sweep.txt
