# round 2: full GPU suite + bench lines (driver-style 20 steps, and 1000 steps) + reference arm
tag=${1:-r2h}
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader; nproc
echo "=== pytest"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "=== smoke"; python __graft_entry__.py smoke 2>&1 | tail -4
echo "=== bench 20"; python bench.py --steps 20 --warmup 3 | tee gpurun_out/bench20_$tag.json
echo "=== bench 1000"; python bench.py --steps 1000 --warmup 10 --no-cpu | tee gpurun_out/bench1000_$tag.json
echo "=== reference arm"; python bench.py --impl reference --steps 20 --warmup 3
