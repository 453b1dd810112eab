mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q > gpurun_out/k9.log 2>&1; tail -3 gpurun_out/k9.log
timeout 900 python -m pytest tests/test_unet_gpu.py -q > gpurun_out/u6.log 2>&1; tail -3 gpurun_out/u6.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench6.json 2> gpurun_out/bench6.err; tail -2 gpurun_out/bench6.err
python -c "
import json; d=json.load(open('gpurun_out/bench6.json')); print('ms/image', d['value'], 'attn TF/s', d['roofline']['achieved'], 'gn ms', d['roofline']['groupnorm']['ms_per_image'], 'launches', d['gpu_launches'], d['clocks'])"
