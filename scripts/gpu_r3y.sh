# midpoint: Newton reciprocal in the Cayley transform instead of the FP64 division
mkdir -p gpurun_out
exec > gpurun_out/r3y.log 2>&1
AB_REPS=2 python scripts/abbench.py
AB_TEMP=0 AB_REPS=1 python scripts/abbench.py
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
