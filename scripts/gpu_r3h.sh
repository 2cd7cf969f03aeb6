# secondary workloads on the new build: fcc z = 18 (staged kernel by default; run kernel forced), config 4 (skybench), Monte Carlo
mkdir -p gpurun_out
exec > gpurun_out/r3h.log 2>&1
echo "=== fcc default"; python scripts/fccbench.py
echo "=== fcc, run kernel on 1024-slot tiles (ASD_RUNS=1024)"; ASD_RUNS=1024 python scripts/fccbench.py
echo "=== fcc, run kernel, no planes"; ASD_RUNS=1024 ASD_MM=0 python scripts/fccbench.py
echo "=== config 4"; python scripts/skybench.py
echo "=== mc"; python scripts/mcbench.py
