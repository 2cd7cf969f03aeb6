# round 2, closing run: full GPU suite, smoke, the driver-style bench line, reference arm, launch list of the bench command
tag=${1:-r2o}
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader; nproc
echo "=== pytest"; timeout 2400 python -m pytest tests -m gpu -q -W ignore 2>&1 | tail -6
echo "=== smoke"; python __graft_entry__.py smoke 2>&1 | tail -4
echo "=== bench 20"; python bench.py --steps 20 --warmup 3 > gpurun_out/bench20_$tag.json; cut -c1-400 gpurun_out/bench20_$tag.json
echo "=== reference arm"; python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/ref20_$tag.json; cut -c1-300 gpurun_out/ref20_$tag.json
echo "=== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-secondary > gpurun_out/ncul_$tag.log 2>&1; tail -2 gpurun_out/ncul_$tag.log | cut -c1-200
