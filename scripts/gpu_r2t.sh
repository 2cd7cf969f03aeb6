# round 2: ncu --set full of the staged one-atom-per-thread stage kernel on fcc z = 18 (128 x 128 x 64 x 4 atoms)
tag=${1:-r2t}
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:llg_stage -s 12 -c 2 -o gpurun_out/prof_$tag python scripts/fccbench.py > gpurun_out/ncu_$tag.log 2>&1
tail -3 gpurun_out/ncu_$tag.log
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > gpurun_out/prof_${tag}_source.csv
python scripts/ncu_summary.py gpurun_out/prof_${tag}_raw.csv
