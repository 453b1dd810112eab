# round 2, GPU call U (1 GPU): SD1.x heads stored 64 wide (zero-padded projections) -- parity on the SD1.x cases, A/B bench
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -k "sd15 or w2_nosplit" --durations=3 > gpurun_out/r2u_tests.log 2>&1; echo "tests rc=$?"; tail -7 gpurun_out/r2u_tests.log
for P in 1 0; do
  DF_PAD_HEADS=$P timeout 200 python bench.py --model sd15 --steps 3 --warmup 3 --no-roofline --no-cpu-baseline > gpurun_out/r2u_bench_sd15_pad$P.json 2> gpurun_out/r2u_bench_sd15_pad$P.err; tail -1 gpurun_out/r2u_bench_sd15_pad$P.err | cut -c1-200; cut -c1-220 gpurun_out/r2u_bench_sd15_pad$P.json
done
