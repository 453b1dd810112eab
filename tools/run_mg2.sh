mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_gpu.py -q -k "w2_nosplit or w2_stale or w2_fullsync or sd15" > gpurun_out/mg2_tests.log 2>&1; tail -3 gpurun_out/mg2_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-split-batch > gpurun_out/bench_n2_nosplit.json 2> gpurun_out/bench_n2_nosplit.err; tail -2 gpurun_out/bench_n2_nosplit.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/bench_n2_nosplit.json')); print('N=2 patch2 ms/image', d['value'], 'exposed', d['exposed_comm'], 'attn', d['roofline']['achieved'])"
