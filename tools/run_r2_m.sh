# round 2, GPU call M (2 GPUs, patch split): does a high-priority compute stream keep the publication kernels from displacing
# persistent attention CTAs?  1024^2 and 2048^2 (hires block), exposed communication split by step kind
mkdir -p gpurun_out
run() { name=$1; port=$2; shift; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 2 --warmup 3 --no-split-batch --hires-resolution 2048 --no-roofline > gpurun_out/r2m_$name.json 2> gpurun_out/r2m_$name.err; }
run prio0 29531 DF_COMPUTE_PRIO=0
run prio_hi 29532 DF_COMPUTE_PRIO=-1
run prio0_b 29533 DF_COMPUTE_PRIO=0
run prio_hi_b 29534 DF_COMPUTE_PRIO=-1
python - <<'PY'
import json
for n in ("prio0", "prio_hi", "prio0_b", "prio_hi_b"):
    try:
        d = json.load(open(f"gpurun_out/r2m_{n}.json")); e = d["exposed_comm"]; h = d["hires"]; he = h.get("exposed_comm", {})
        print(f"{n:10s} 1024: {d['value']:7.1f} ms exposed {e['exposed_comm_pct']:.2f}% (sync {e['sync_step_ms']:.2f} async {e['async_step_ms']:.2f} compute {e['compute_only_step_ms']:.2f})"
              f" | 2048: {h['ms_per_image']:8.1f} ms sync {he.get('sync_step_ms', 0):.2f} async {he.get('async_step_ms', 0):.2f} compute {he.get('compute_only_step_ms', 0):.2f} exposed/async {he.get('exposed_pct_async_step', 0):.2f}% image {he.get('exposed_pct_image_from_steps', 0):.2f}%")
    except Exception as ex:
        print(n, "failed", ex)
PY
