mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_unet_gpu.py -q -k "multi_rank and (w2_nosplit or w2_stale or cuda_graph)" > gpurun_out/mg2_tests.log 2>&1; tail -5 gpurun_out/mg2_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -5 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-split-batch > gpurun_out/bench_n2_nosplit.json 2> gpurun_out/bench_n2_nosplit.err; tail -5 gpurun_out/bench_n2_nosplit.err; cat gpurun_out/bench_n2_nosplit.json
