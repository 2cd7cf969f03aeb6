# closing check of the final tree + L2 prefetch distance of the moment-plane kernel (ASD_PF, tiles ahead)
mkdir -p gpurun_out
exec > gpurun_out/r3z.log 2>&1
echo "=== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for pf in 0 74 148 222 296 444; do echo "ASD_PF=$pf"; ASD_PF=$pf AB_REPS=1 python scripts/abbench.py; done
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -2
