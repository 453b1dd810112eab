# round 2, GPU call Y (1 GPU): scheduler signalling through shared-memory atomics -- attention parity, racecheck, two UNet cases
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_kernels_gpu.py -q -k "attention" > gpurun_out/r2y_attn_tests.log 2>&1; tail -2 gpurun_out/r2y_attn_tests.log
timeout 150 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_kernels_gpu.py -q -x -k "single_segment and (256-384 or 200-77-2-64) or large_logits_rescale and (x40 or some_rows) or groupnorm_local and 320 or fused_halo and True or bias_residual_add and 320" > gpurun_out/r2y_racecheck.log 2>&1; tail -4 gpurun_out/r2y_racecheck.log
timeout 150 python -m pytest tests/test_unet_gpu.py -q -k "w2_nosplit or full_size" > gpurun_out/r2y_unet.log 2>&1; tail -2 gpurun_out/r2y_unet.log
timeout 60 python tools/bench_attn.py --shapes 1024_l1,3840n4_l2 2>&1 | tail -2
