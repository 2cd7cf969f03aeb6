# full suite + bench on the current build
mkdir -p gpurun_out
exec > gpurun_out/r3q.log 2>&1
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "=== bench 20"; python bench.py --steps 20 --warmup 3 | tee gpurun_out/bench20_r3q.json | cut -c1-300
