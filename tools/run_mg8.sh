mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/ngpu.txt
N=${1:-8}
for R in 1024 3840; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$((R/1000)) bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline --resolution $R > gpurun_out/bench_n${N}_${R}.json 2> gpurun_out/bench_n${N}_${R}.err
tail -3 gpurun_out/bench_n${N}_${R}.err | cut -c1-300; cut -c1-2600 gpurun_out/bench_n${N}_${R}.json
done
