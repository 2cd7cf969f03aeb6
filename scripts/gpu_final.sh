# usage: bash scripts/gpu_final.sh <tag>  -- tests, bench, launch list, full ncu capture of the stage kernels
tag=${1:-final}
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
echo "=== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "=== smoke"; python __graft_entry__.py smoke 2>&1 | tail -3
echo "=== bench"; python bench.py --steps 1000 --warmup 10 | tee gpurun_out/bench_$tag.json
echo "=== bench T=0"; python bench.py --steps 500 --warmup 10 --temp 0 --no-cpu
echo "=== bench depondt"; python bench.py --steps 500 --warmup 10 --solver 5 --no-cpu
echo "=== bench old kernel (ASD_RUNS=0)"; ASD_RUNS=0 python bench.py --steps 500 --warmup 10 --no-cpu
echo "=== reference arm"; python bench.py --impl reference --steps 5 --warmup 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:llg_ -s 6 -c 2 -o gpurun_out/prof_$tag python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_$tag.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncul_$tag.log 2>&1
