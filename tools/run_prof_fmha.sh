mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err; tail -3 gpurun_out/bench2.err; cat gpurun_out/bench2.json
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:fmha_fwd -s 0 -c 2 -f -o gpurun_out/r1_fmha_lvl1 python tools/profile_step.py --resolution 1024 > gpurun_out/prof2.log 2>&1; tail -2 gpurun_out/prof2.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:fmha_fwd -s 8 -c 1 -f -o gpurun_out/r1_fmha_lvl2 python tools/profile_step.py --resolution 1024 > gpurun_out/prof3.log 2>&1; tail -2 gpurun_out/prof3.log
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:gn_ -s 0 -c 3 -f -o gpurun_out/r1_gn_320 python tools/profile_step.py --resolution 1024 > gpurun_out/prof4.log 2>&1; tail -2 gpurun_out/prof4.log
ls -la gpurun_out/*.ncu-rep
