tag=${1:-mc1}
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== mcbench"; python scripts/mcbench.py 64 64 64; python scripts/mcbench.py 128 128 128
