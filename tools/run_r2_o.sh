# round 2, GPU call O (1 GPU): attention register / polynomial-share variants under setmaxnreg; op-level view of one eager step
mkdir -p gpurun_out
SH="1024_l1,1024_l2,3840n4_l2,3840n4_l1"
rm -f gpurun_out/r2o_attn_sweep.txt
DF_LIB_PATH=distrifuser_b200/variants/lib_wg10.so timeout 90 python -m pytest tests/test_kernels_gpu.py -q -x -k "single_segment and 256-384" > gpurun_out/r2o_wg10_test.log 2>&1; echo "wg10 test rc=$?"; tail -3 gpurun_out/r2o_wg10_test.log
for V in main wg10 smr_emu3 smr_emu5 smr_emu6 main; do
  echo "== variant: $V" >> gpurun_out/r2o_attn_sweep.txt
  if [ $V = main ]; then timeout 120 python tools/bench_attn.py --shapes $SH >> gpurun_out/r2o_attn_sweep.txt 2>&1; else DF_LIB_PATH=distrifuser_b200/variants/lib_$V.so timeout 120 python tools/bench_attn.py --shapes $SH >> gpurun_out/r2o_attn_sweep.txt 2>&1; fi
done
cat gpurun_out/r2o_attn_sweep.txt
timeout 200 python tools/profile_step_ops.py > gpurun_out/r2o_step_ops.txt 2>&1; head -60 gpurun_out/r2o_step_ops.txt
