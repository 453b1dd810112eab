# round 2, GPU call L (1 GPU): GroupNorm with per-CTA statistics (variant gn2) -- parity, graph-timed A/B, phase trace
mkdir -p gpurun_out
DF_LIB_PATH=distrifuser_b200/variants/lib_gn2.so timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "groupnorm or halo" > gpurun_out/r2l_gn_tests.log 2>&1; tail -4 gpurun_out/r2l_gn_tests.log
DF_LIB_PATH=distrifuser_b200/variants/lib_gn2.so timeout 400 python -m pytest tests/test_unet_gpu.py -q -x -k "w2_nosplit or w2_stale or w2_syncgn or sd15_multi" > gpurun_out/r2l_unet_tests.log 2>&1; tail -4 gpurun_out/r2l_unet_tests.log
for V in main gn2 main gn2; do
  echo "== $V" >> gpurun_out/r2l_gn_bench.txt
  if [ $V = main ]; then timeout 120 python tools/bench_gn.py >> gpurun_out/r2l_gn_bench.txt 2>&1; else DF_LIB_PATH=distrifuser_b200/variants/lib_$V.so timeout 120 python tools/bench_gn.py >> gpurun_out/r2l_gn_bench.txt 2>&1; fi
done
cat gpurun_out/r2l_gn_bench.txt
DF_LIB_PATH=distrifuser_b200/variants/lib_gn2trace.so timeout 120 python tools/trace_gn.py > gpurun_out/r2l_gn_trace.txt 2>&1; cat gpurun_out/r2l_gn_trace.txt
