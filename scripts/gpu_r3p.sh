# Monte Carlo block sweep with pre-drawn trial moves (mc_predraw_kernel) against draws inside the sweep CTAs; MC tests
mkdir -p gpurun_out
exec > gpurun_out/r3p.log 2>&1
echo "=== predraw (default)"; python scripts/mcbench.py 128 128 128
echo "=== draws inside the sweep kernel (ASD_MC_PREDRAW=0)"; ASD_MC_PREDRAW=0 python scripts/mcbench.py 128 128 128
echo "=== pytest MC"; timeout 1500 python -m pytest tests/test_gpu_mc_parity.py tests/test_gpu_mc_configs.py tests/test_gpu_alloy.py -m gpu -x -q 2>&1 | tail -4
