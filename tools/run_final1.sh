mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/final_gpu_tests.log 2>&1; tail -12 gpurun_out/final_gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -2 gpurun_out/bench_final_n1.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/bench_final_n1.json')); print('ms/image', d['value'], 'e2e', d['e2e']['value'], 'attn', d['roofline']['achieved'], d['roofline']['frac'], 'launches', d['gpu_launches'], d['clocks'])"
