# lean integrator in the XS kernels (config 4), fcc on the run kernel by default, suite
mkdir -p gpurun_out
exec > gpurun_out/r3i.log 2>&1
echo "=== config 4: lean integrator / general"; python scripts/skybench.py; ASD_LEAN=0 python scripts/skybench.py
echo "=== fcc default"; python scripts/fccbench.py
echo "=== headline"; AB_REPS=1 python scripts/abbench.py
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
