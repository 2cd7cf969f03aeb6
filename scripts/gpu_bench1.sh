tag=${1:-b1}
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "=== bench N=1"; timeout 900 python bench.py --steps 200 --warmup 10
echo "=== bench N=1 slab (self halo)"; timeout 900 python bench.py --steps 200 --warmup 10 --decomp slab --no-cpu
echo "=== bench N=1 depondt"; timeout 900 python bench.py --steps 200 --warmup 10 --solver 5 --no-cpu
echo "=== bench reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 3
