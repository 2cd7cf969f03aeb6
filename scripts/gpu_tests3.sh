# usage: bash scripts/gpu_tests3.sh <tag>  -- GPU parity suite + smoke + headline bench (no CPU baseline)
tag=${1:-tests}
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30
echo "=== smoke"; python __graft_entry__.py smoke 2>&1 | tail -5
echo "=== bench"; python bench.py --steps 500 --warmup 10 --no-cpu | cut -c1-330
echo "=== bench T=0"; python bench.py --steps 500 --warmup 10 --temp 0 --no-cpu | cut -c1-330
