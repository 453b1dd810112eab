# round 2, GPU call K (1 GPU): what holding S costs (unsafe early-release build, timing only), ncu of the speculative-reference
# kernel, GroupNorm phase trace
mkdir -p gpurun_out
SH="1024_l1,1024_l2,3840n4_l2,3840n4_l1"
rm -f gpurun_out/r2k_attn_sweep.txt
for V in emu4 early_unsafe emu4 early_unsafe; do
  echo "== variant: $V" >> gpurun_out/r2k_attn_sweep.txt
  DF_LIB_PATH=distrifuser_b200/variants/lib_$V.so timeout 200 python tools/bench_attn.py --shapes $SH >> gpurun_out/r2k_attn_sweep.txt 2>&1
done
cat gpurun_out/r2k_attn_sweep.txt
DF_LIB_PATH=distrifuser_b200/variants/lib_gntrace.so timeout 120 python tools/trace_gn.py > gpurun_out/r2k_gn_trace.txt 2>&1; cat gpurun_out/r2k_gn_trace.txt
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/r2k_attn python tools/ncu_kernels.py attn3840 > gpurun_out/r2k_ncu.log 2>&1; tail -2 gpurun_out/r2k_ncu.log
