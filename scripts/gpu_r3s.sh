# final ncu --set full of the MM + LEAN run kernels (padded union rows, pairs walk) + launch list + bench line + reference arm
tag=r3s
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:llg_runs -s 6 -c 2 -o gpurun_out/prof_$tag python bench.py --steps 3 --warmup 3 --no-cpu --no-secondary > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log | cut -c1-200
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > gpurun_out/prof_${tag}_source.csv
python scripts/ncu_summary.py gpurun_out/prof_${tag}_raw.csv
echo "=== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-secondary > gpurun_out/ncul_$tag.log 2>&1
echo "=== bench 20"; python bench.py --steps 20 --warmup 3 | tee gpurun_out/bench20_$tag.json | cut -c1-200
echo "=== reference arm"; python bench.py --impl reference --steps 20 --warmup 3 | tee gpurun_out/ref20_$tag.json | cut -c1-400
