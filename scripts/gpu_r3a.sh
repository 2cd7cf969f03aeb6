# ablation A/B of the run kernel + LEAN instantiation (round 2, session 4)
mkdir -p gpurun_out
exec > gpurun_out/r3a.log 2>&1
echo "=== default build (LEAN on / off), midpoint 300 K"
AB_REPS=2 python scripts/abbench.py
ASD_LEAN=0 AB_REPS=2 python scripts/abbench.py
echo "=== T = 0"
AB_TEMP=0 AB_REPS=1 python scripts/abbench.py
ASD_LEAN=0 AB_TEMP=0 AB_REPS=1 python scripts/abbench.py
echo "=== Depondt"
AB_SOLVER=5 AB_REPS=1 python scripts/abbench.py
ASD_LEAN=0 AB_SOLVER=5 AB_REPS=1 python scripts/abbench.py
echo "=== ablations (results wrong by construction): abl1 no integrator math, abl2 no union walk, abl4 no staging, abl7 none of them; unr2: RUN_UNROLL=2"
AB_REPS=1 python scripts/abbench.py build_var/abl1.so build_var/abl2.so build_var/abl4.so build_var/abl7.so build_var/unr2.so
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
