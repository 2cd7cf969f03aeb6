mkdir -p gpurun_out
exec > gpurun_out/mcncu.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/kbench.py
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mc_colour -s 40 -c 3 -o gpurun_out/prof_mc python scripts/mcbench.py 128 128 128 > gpurun_out/ncu_mc.log 2>&1
tail -5 gpurun_out/ncu_mc.log
