# usage: bash scripts/gpu_tests2.sh <tag>  -- GPU parity suite + smoke + small-system bench
tag=${1:-tests}
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30
echo "=== smoke"; python __graft_entry__.py smoke 2>&1 | tail -5
echo "=== smallbench"; timeout 600 python scripts/smallbench.py 2>&1 | tail -30
