mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m pytest tests/test_unet_gpu.py -q -k "eight_gpus" > gpurun_out/mg8_tests.log 2>&1; tail -3 gpurun_out/mg8_tests.log
for R in 1024 3840; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$((R/1000)) bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline --resolution $R > gpurun_out/bench_n${N}_${R}.json 2> gpurun_out/bench_n${N}_${R}.err
grep -i "error\|traceback" gpurun_out/bench_n${N}_${R}.err | head -3; cut -c1-2700 gpurun_out/bench_n${N}_${R}.json
done
