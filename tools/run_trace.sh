mkdir -p gpurun_out
DF_NVCC_FLAGS="-DDF_TRACE" python -c "from distrifuser_b200 import build; build.build(force=True)" > /dev/null
python tools/trace_attn.py 1 3600 14400 20 64 > gpurun_out/trace_big.txt 2>&1; head -30 gpurun_out/trace_big.txt
python tools/trace_attn.py 1 256 1024 20 64 > gpurun_out/trace_small.txt 2>&1; head -12 gpurun_out/trace_small.txt
