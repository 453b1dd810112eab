# round 2, GPU call G (2 GPUs): exposed communication with the slim / deep publication kernel and the normal-priority comm stream
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 python -m pytest tests/test_unet_gpu.py -q -k "w2_nosplit or w2_stale or w2_syncgn" > gpurun_out/r2g_tests.log 2>&1; tail -3 gpurun_out/r2g_tests.log
run() { name=$1; shift; env "$@" timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-split-batch --no-hires --no-roofline > gpurun_out/r2g_$name.json 2> gpurun_out/r2g_$name.err; }
run default DF_NOTHING=1
run prio_high DF_COMM_PRIO=-1
run ctas64 DF_PUB_CTAS=64
run ctas8 DF_PUB_CTAS=8
python - <<'PY'
import json
for n in ("default", "prio_high", "ctas64", "ctas8"):
    try:
        d = json.load(open(f"gpurun_out/r2g_{n}.json")); e = d["exposed_comm"]
        print(f"{n:10s} {d['value']:7.1f} ms  exposed {e['exposed_comm_pct']:.2f}%  sync {e['sync_step_ms']:.2f}  async {e['async_step_ms']:.2f}  compute {e['compute_only_step_ms']:.2f}  ex_sync {e['exposed_ms_per_sync_step']:.2f}  ex_async {e['exposed_ms_per_async_step']:.2f}")
    except Exception as ex:
        print(n, "failed", ex)
PY
