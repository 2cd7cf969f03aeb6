mkdir -p gpurun_out
exec > gpurun_out/one.log 2>&1
python scripts/layoutprobe.py
