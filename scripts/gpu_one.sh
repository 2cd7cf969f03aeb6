mkdir -p gpurun_out
exec > gpurun_out/one.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
AB_REPS=1 python scripts/abbench.py
timeout 900 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -2
