# round 2: ncu --set full of the Monte Carlo block sweep kernel (bcc 128^3, Metropolis + heat bath)
tag=${1:-r2b}
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mc_block_run_kernel -s 6 -c 1 -o gpurun_out/prof_$tag python scripts/mcbench.py 128 128 128 > gpurun_out/ncu_$tag.log 2>&1
tail -5 gpurun_out/ncu_$tag.log
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > gpurun_out/prof_${tag}_source.csv
python scripts/ncu_summary.py gpurun_out/prof_${tag}_raw.csv
