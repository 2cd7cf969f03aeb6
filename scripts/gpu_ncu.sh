tag=${1:-p1}
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
python scripts/kbench.py
timeout 900 ncu --set full --clock-control none --import-source on -k regex:llg_stage -s 6 -c 2 -o gpurun_out/prof_$tag python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_$tag.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncul_$tag.log 2>&1
