# round 2, GPU call I (8 GPUs): 8-rank parity tests, the headline bench at N=8 (1024^2 + the 3840^2 hires block with the exposed-
# communication split)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > gpurun_out/r2i_gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 330 $TR bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2i_bench_n8.json 2> gpurun_out/r2i_bench_n8.err; tail -3 gpurun_out/r2i_bench_n8.err; cut -c1-300 gpurun_out/r2i_bench_n8.json
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2i_bench_n8.json"))
    print("N=8 1024:", d["value"], "ms; exposed", (d.get("exposed_comm") or {}).get("exposed_comm_pct"), "%")
    print("hires:", json.dumps(d.get("hires"))[:900])
except Exception as e:
    print("bench failed", e)
PY
timeout 200 python -m pytest tests/test_unet_gpu.py -q -m gpu -k "eight" --durations=4 > gpurun_out/r2i_tests.log 2>&1; tail -8 gpurun_out/r2i_tests.log
