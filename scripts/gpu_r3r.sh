# lean integrator in the staged one-atom-per-thread kernel (layouts without the run regularity; forced here with ASD_RUNS=0)
mkdir -p gpurun_out
exec > gpurun_out/r3r.log 2>&1
echo "=== staged kernel, lean integrator"; ASD_RUNS=0 AB_REPS=1 python scripts/abbench.py
echo "=== staged kernel, general integrator"; ASD_RUNS=0 ASD_LEAN=0 AB_REPS=1 python scripts/abbench.py
echo "=== Depondt"; ASD_RUNS=0 AB_SOLVER=5 AB_REPS=1 python scripts/abbench.py; ASD_RUNS=0 ASD_LEAN=0 AB_SOLVER=5 AB_REPS=1 python scripts/abbench.py
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
