# slabs on the moment planes: slab tests on one GPU, then the whole suite
mkdir -p gpurun_out
exec > gpurun_out/r3k.log 2>&1
echo "=== slab tests"; timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -15
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "=== headline"; AB_REPS=1 python scripts/abbench.py
