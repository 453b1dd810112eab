mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q > gpurun_out/k7.log 2>&1; tail -3 gpurun_out/k7.log
timeout 600 python -m pytest tests/test_unet_gpu.py -q -k "single_gpu or w2_nosplit or w4_split or syncgn" > gpurun_out/u4.log 2>&1; tail -3 gpurun_out/u4.log
python tools/bench_attn.py --shapes 1024_l1,1024_l2,3840n4_l2 2>&1 | tee gpurun_out/attn_v2b.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; tail -2 gpurun_out/bench5.err
python -c "
import json; d=json.load(open('gpurun_out/bench5.json')); print('ms/image', d['value'], 'attn TF/s', d['roofline']['achieved'], 'gn', d['roofline']['groupnorm'], 'launches', d['gpu_launches'])"
