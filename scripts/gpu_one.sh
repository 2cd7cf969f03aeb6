mkdir -p gpurun_out
exec > gpurun_out/one.log 2>&1
timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -4
