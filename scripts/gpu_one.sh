mkdir -p gpurun_out
exec > gpurun_out/one.log 2>&1
ASD_DEBUG=1 python scripts/layoutprobe.py 100x8x8 48x8x8 40x8x8 33x8x8 2>&1 | grep -v "^$" | tail -30
