# round 2, GPU call T (1 GPU): why is the d=40 (SD1.x level 0) attention shape at half the per-tile rate of d=64?
mkdir -p gpurun_out
timeout 120 python tools/bench_attn.py --shapes sd15_l0,sd15_l1,3840n4_l2 > gpurun_out/r2t_attn.txt 2>&1; cat gpurun_out/r2t_attn.txt
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/r2t_sd15 python tools/bench_attn.py --shapes sd15_l0 --profile > gpurun_out/r2t_ncu.log 2>&1; tail -2 gpurun_out/r2t_ncu.log
