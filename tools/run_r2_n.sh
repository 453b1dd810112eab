# round 2, GPU call N (1 GPU): attention with the item ring (dynamic tickets), early S release, in-register repair and replay --
# parity tests (wrapped in a timeout: a scheduling bug would hang), then graph-timed A/B against the hold-S build
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" > gpurun_out/r2n_attn_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2n_attn_tests.log
SH="1024_l1,1024_l2,3840n4_l2,3840n4_l1,1024n4_l1"
rm -f gpurun_out/r2n_attn_sweep.txt
for V in hold main v5c_nosmr early_unsafe; do
  echo "== variant: $V" >> gpurun_out/r2n_attn_sweep.txt
  if [ $V = main ]; then timeout 200 python tools/bench_attn.py --shapes $SH >> gpurun_out/r2n_attn_sweep.txt 2>&1; else DF_LIB_PATH=distrifuser_b200/variants/lib_$V.so timeout 200 python tools/bench_attn.py --shapes $SH >> gpurun_out/r2n_attn_sweep.txt 2>&1; fi
done
cat gpurun_out/r2n_attn_sweep.txt
timeout 200 python tools/profile_step_ops.py > gpurun_out/r2n_step_ops.txt 2>&1; head -50 gpurun_out/r2n_step_ops.txt
