mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1_1024.json 2> gpurun_out/bench_n1_1024.err; tail -2 gpurun_out/bench_n1_1024.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n1_1024.json')); print('1024: ms/image', d['value'], 'e2e', d['e2e']['value'], 'attn', d['roofline']['achieved'], d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'], d['clocks'])"
timeout 900 python bench.py --steps 1 --warmup 3 --resolution 3840 --no-cpu-baseline > gpurun_out/bench_n1_3840.json 2> gpurun_out/bench_n1_3840.err; tail -2 gpurun_out/bench_n1_3840.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n1_3840.json')); print('3840: ms/image', d['value'], 'attn', d['roofline']['achieved'], d['roofline']['frac'], d['clocks'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -2 gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/bench_ref.json
