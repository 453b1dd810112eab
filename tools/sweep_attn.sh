mkdir -p gpurun_out
for E in 0 2 3 4; do
  DF_NVCC_FLAGS="-DDF_EMU_PAIRS_OF_8=$E" python -c "from distrifuser_b200 import build; build.build(force=True)" > /dev/null
  echo "== EMU_PAIRS_OF_8=$E" >> gpurun_out/attn_sweep.txt
  python tools/bench_attn.py --shapes 1024_l1,1024_l2,3840n4_l2 >> gpurun_out/attn_sweep.txt 2>&1
done
cat gpurun_out/attn_sweep.txt
