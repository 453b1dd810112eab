# round 2, GPU call S (8 GPUs): the headline configuration with the final build (1024^2 + 3840^2 hires block, exposed communication)
mkdir -p gpurun_out
timeout 190 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 1 --warmup 3 --no-roofline > gpurun_out/r2s_bench_n8.json 2> gpurun_out/r2s_bench_n8.err; tail -2 gpurun_out/r2s_bench_n8.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2s_bench_n8.json")); e = d["exposed_comm"]; h = d["hires"]; he = h.get("exposed_comm", {})
    print(f"N=8 1024: {d['value']:7.1f} ms exposed {e['exposed_comm_pct']:.2f}% (sync {e['sync_step_ms']:.2f} async {e['async_step_ms']:.2f} compute {e['compute_only_step_ms']:.2f})"
          f" | 3840: {h['ms_per_image']:8.1f} ms sync {he.get('sync_step_ms', 0):.2f} async {he.get('async_step_ms', 0):.2f} compute {he.get('compute_only_step_ms', 0):.2f} exposed/async {he.get('exposed_pct_async_step', 0):.2f}% image {he.get('exposed_pct_image_from_steps', 0):.2f}%")
except Exception as ex:
    print("failed", ex)
PY
