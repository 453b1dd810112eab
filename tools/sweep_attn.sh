mkdir -p gpurun_out; rm -f gpurun_out/attn_sweep.txt
for V in "-DDF_TRYWAIT_HINT_NS=20000u" "-DDF_TRYWAIT_HINT_NS=1000u" "-DDF_TRYWAIT_HINT_NS=200000u -DDF_SPIN_FAST_POLLS=1" ; do
  DF_NVCC_FLAGS="$V" python -c "from distrifuser_b200 import build; build.build(force=True)" > /dev/null
  echo "== $V" >> gpurun_out/attn_sweep.txt
  python tools/bench_attn.py --shapes 1024_l1,1024_l2,3840n4_l2 >> gpurun_out/attn_sweep.txt 2>&1
done
python -c "from distrifuser_b200 import build; build.build(force=True)" > /dev/null
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "attention" > gpurun_out/k8.log 2>&1; tail -4 gpurun_out/k8.log
timeout 300 python -m pytest tests/test_unet_gpu.py -q -k "sd15" > gpurun_out/u5.log 2>&1; tail -4 gpurun_out/u5.log
cat gpurun_out/attn_sweep.txt
