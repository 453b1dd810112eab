# round 2, GPU call E (2 GPUs): real-NVLink parity of the world-2 cases, patch2 bench with the exposed-comm split, fused q|k|v publication A/B
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2e_gpus.txt
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_pipeline_gpu.py -q -k "w2 or sd15_multi or multi_rank_cuda_graph" > gpurun_out/r2e_tests.log 2>&1; tail -5 gpurun_out/r2e_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-split-batch --no-hires > gpurun_out/r2e_bench_n2_patch2.json 2> gpurun_out/r2e_bench_n2_patch2.err; tail -2 gpurun_out/r2e_bench_n2_patch2.err
DF_LINEAR=geglu,qkv timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-split-batch --no-hires --no-roofline > gpurun_out/r2e_bench_n2_patch2_qkv.json 2> gpurun_out/r2e_bench_n2_patch2_qkv.err; tail -2 gpurun_out/r2e_bench_n2_patch2_qkv.err
DF_FUSED_HALO=0 timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-split-batch --no-hires --no-roofline > gpurun_out/r2e_bench_n2_patch2_nohalofuse.json 2> gpurun_out/r2e_bench_n2_patch2_nohalofuse.err
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-hires --no-roofline > gpurun_out/r2e_bench_n2_cfg.json 2> gpurun_out/r2e_bench_n2_cfg.err
python - <<'PY'
import json
for n in ("n2_patch2", "n2_patch2_qkv", "n2_patch2_nohalofuse", "n2_cfg"):
    try:
        d = json.load(open(f"gpurun_out/r2e_bench_{n}.json"))
        print(n, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "exposed", json.dumps(d.get("exposed_comm"))[:600])
    except Exception as e:
        print(n, "failed", e)
PY
