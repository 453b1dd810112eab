mkdir -p gpurun_out; rm -f gpurun_out/attn_sweep.txt
for V in "-DDF2_ORDER_EXP=0" "-DDF2_ORDER_EXP=0 -DDF2_EMU_PAIRS_OF_8=0" "-DDF2_ORDER_EXP=0 -DDF2_EMU_PAIRS_OF_8=3" "-DDF2_ORDER_EXP=1 -DDF2_EMU_PAIRS_OF_8=0"; do
  DF_NVCC_FLAGS="$V" python -c "from distrifuser_b200 import build; build.build(force=True)" > /dev/null
  echo "== $V" >> gpurun_out/attn_sweep.txt
  timeout 120 python tools/bench_attn.py --shapes 1024_l1,3840n4_l2 --kernel 2 >> gpurun_out/attn_sweep.txt 2>&1
done
cat gpurun_out/attn_sweep.txt
