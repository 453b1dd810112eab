# round 2, GPU call P (1 GPU): full GPU suite on the current build (item-ring attention, per-CTA GroupNorm statistics, conv bias /
# residual pass), default bench, attention variants
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2p_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2p_smoke.log 2>&1; tail -2 gpurun_out/r2p_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err; tail -2 gpurun_out/r2p_bench_n1.err; cut -c1-300 gpurun_out/r2p_bench_n1.json
DF_CONV_BIAS=torch timeout 300 python bench.py --steps 3 --warmup 3 --no-hires --no-cpu-baseline --no-roofline > gpurun_out/r2p_bench_n1_torchbias.json 2> gpurun_out/r2p_bench_n1_torchbias.err; cut -c1-200 gpurun_out/r2p_bench_n1_torchbias.json
SH="1024_l1,1024_l2,3840n4_l2,3840n4_l1"
rm -f gpurun_out/r2p_attn_sweep.txt
DF_LIB_PATH=distrifuser_b200/variants/lib_wg10.so timeout 90 python -m pytest tests/test_kernels_gpu.py -q -x -k "single_segment and 256-384" > gpurun_out/r2p_wg10_test.log 2>&1; echo "wg10 test rc=$?"; tail -3 gpurun_out/r2p_wg10_test.log
for V in main wg10 smr_emu3 smr_emu5 nosmr; do
  echo "== variant: $V" >> gpurun_out/r2p_attn_sweep.txt
  if [ $V = main ]; then timeout 120 python tools/bench_attn.py --shapes $SH >> gpurun_out/r2p_attn_sweep.txt 2>&1; else DF_LIB_PATH=distrifuser_b200/variants/lib_$V.so timeout 120 python tools/bench_attn.py --shapes $SH >> gpurun_out/r2p_attn_sweep.txt 2>&1; fi
done
cat gpurun_out/r2p_attn_sweep.txt
