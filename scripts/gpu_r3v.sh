# Monte Carlo: reciprocals instead of repeated FP64 divisions in the heat-bath frame and the Gaussian trial move
mkdir -p gpurun_out
exec > gpurun_out/r3v.log 2>&1
python scripts/mcbench.py 128 128 128
python scripts/skybench.py 2>&1 | grep "heat bath"
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
