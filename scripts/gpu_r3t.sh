# union rows without null entries (odd classes hand an entry to the class of their union, zero coupling) against null padding
mkdir -p gpurun_out
exec > gpurun_out/r3t.log 2>&1
echo "=== promoted (default)"; AB_REPS=2 python scripts/abbench.py
echo "=== null entries (ASD_WALK_PROMOTE=0)"; ASD_WALK_PROMOTE=0 AB_REPS=2 python scripts/abbench.py
echo "=== T=0 / Depondt promoted"; AB_TEMP=0 AB_REPS=1 python scripts/abbench.py; AB_SOLVER=5 AB_REPS=1 python scripts/abbench.py
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
