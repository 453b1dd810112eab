# round 2, GPU call Q (4 GPUs, one patch group of 4 = the patch layout of the 8-GPU headline run, twice the per-GPU batch):
# exposed communication at 1024^2 and 3840^2 with the high-priority compute stream + ticket-scheduled attention, and with the
# priorities equal for comparison
mkdir -p gpurun_out
run() { name=$1; port=$2; shift; shift; env "$@" timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 4 --steps 2 --warmup 3 --no-split-batch --no-roofline > gpurun_out/r2q_$name.json 2> gpurun_out/r2q_$name.err; tail -2 gpurun_out/r2q_$name.err | cut -c1-300; }
run default 29541 DF_NOTHING=1
run prio0 29542 DF_COMPUTE_PRIO=0
python - <<'PY'
import json
for n in ("default", "prio0"):
    try:
        d = json.load(open(f"gpurun_out/r2q_{n}.json")); e = d["exposed_comm"]; h = d["hires"]; he = h.get("exposed_comm", {})
        print(f"{n:8s} 1024: {d['value']:7.1f} ms exposed {e['exposed_comm_pct']:.2f}% (sync {e['sync_step_ms']:.2f} async {e['async_step_ms']:.2f} compute {e['compute_only_step_ms']:.2f})"
              f" | 3840: {h['ms_per_image']:8.1f} ms sync {he.get('sync_step_ms', 0):.2f} async {he.get('async_step_ms', 0):.2f} compute {he.get('compute_only_step_ms', 0):.2f} exposed/async {he.get('exposed_pct_async_step', 0):.2f}% image {he.get('exposed_pct_image_from_steps', 0):.2f}%")
    except Exception as ex:
        print(n, "failed", ex)
PY
