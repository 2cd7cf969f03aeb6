mkdir -p gpurun_out
exec > gpurun_out/one.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
AB_REPS=1 python scripts/abbench.py; AB_SOLVER=5 AB_REPS=1 python scripts/abbench.py
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -2
