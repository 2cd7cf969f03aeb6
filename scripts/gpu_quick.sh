# usage: bash scripts/gpu_quick.sh <tag> [ncu]   -- parity tests + kernel bench (+ ncu full capture of the stage kernels)
tag=${1:-quick}
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
echo "=== pytest"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== kbench"
python scripts/kbench.py
if [ "$2" = "ncu" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:llg_stage -s 6 -c 2 -o gpurun_out/prof_$tag python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_$tag.log 2>&1
fi
