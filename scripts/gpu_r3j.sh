# the driver's bench command + launch list on the new build
mkdir -p gpurun_out
exec > gpurun_out/r3j.log 2>&1
echo "=== bench 20"; python bench.py --steps 20 --warmup 3 | tee gpurun_out/bench20_r3j.json
echo "=== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r3j.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-secondary > gpurun_out/ncul_r3j.log 2>&1
tail -2 gpurun_out/ncul_r3j.log | cut -c1-300
