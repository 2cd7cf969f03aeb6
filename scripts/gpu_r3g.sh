# noise hand-over predictor -> corrector (default) against recomputing it (ASD_NOISE_STORE=0); suite
mkdir -p gpurun_out
exec > gpurun_out/r3g.log 2>&1
AB_REPS=2 python scripts/abbench.py
ASD_NOISE_STORE=0 AB_REPS=2 python scripts/abbench.py
AB_SOLVER=5 AB_REPS=1 python scripts/abbench.py
ASD_NOISE_STORE=0 AB_SOLVER=5 AB_REPS=1 python scripts/abbench.py
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
