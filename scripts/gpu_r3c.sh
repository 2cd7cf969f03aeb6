# own-spin L1 prefetch A/B + ncu --set full of the MM + LEAN run kernels (round 2, session 4)
tag=r3c
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
echo "=== with / without the own-spin L1 prefetch"
AB_REPS=2 python scripts/abbench.py uppasd_b200/libuppasd_b200.so build_var/nopf.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:llg_runs -s 6 -c 2 -o gpurun_out/prof_$tag python bench.py --steps 3 --warmup 3 --no-cpu --no-secondary > gpurun_out/ncu_$tag.log 2>&1
tail -3 gpurun_out/ncu_$tag.log
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > gpurun_out/prof_${tag}_source.csv
python scripts/ncu_summary.py gpurun_out/prof_${tag}_raw.csv
