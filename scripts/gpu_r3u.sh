# Depondt: one division + three products in the Rodrigues rotation
mkdir -p gpurun_out
exec > gpurun_out/r3u.log 2>&1
AB_SOLVER=5 AB_REPS=2 python scripts/abbench.py
AB_SOLVER=5 AB_TEMP=0 AB_REPS=1 python scripts/abbench.py
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
