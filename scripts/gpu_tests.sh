# usage: bash scripts/gpu_tests.sh <tag>  -- GPU parity suite + smoke only
tag=${1:-tests}
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
echo "=== pytest"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25
echo "=== smoke"; python __graft_entry__.py smoke 2>&1 | tail -4
