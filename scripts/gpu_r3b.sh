# MM (moment planes + cp.async staging) x LEAN A/B, then the GPU suite (round 2, session 4)
mkdir -p gpurun_out
exec > gpurun_out/r3b.log 2>&1
echo "=== MM on, LEAN on"; AB_REPS=2 python scripts/abbench.py
echo "=== MM on, LEAN off"; ASD_LEAN=0 AB_REPS=2 python scripts/abbench.py
echo "=== MM off, LEAN off"; ASD_MM=0 ASD_LEAN=0 AB_REPS=2 python scripts/abbench.py
echo "=== MM off, LEAN on"; ASD_MM=0 AB_REPS=1 python scripts/abbench.py
echo "=== T = 0 (MM on: LEAN on / off)"
AB_TEMP=0 AB_REPS=1 python scripts/abbench.py
ASD_LEAN=0 AB_TEMP=0 AB_REPS=1 python scripts/abbench.py
echo "=== Depondt (MM on: LEAN on / off; MM off)"
AB_SOLVER=5 AB_REPS=1 python scripts/abbench.py
ASD_LEAN=0 AB_SOLVER=5 AB_REPS=1 python scripts/abbench.py
ASD_MM=0 ASD_LEAN=0 AB_SOLVER=5 AB_REPS=1 python scripts/abbench.py
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
