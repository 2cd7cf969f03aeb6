# round 2: warp-specialised MC kernel (ASD_MC_NT=512) against the 256-thread kernel: chain parity + micro-benchmark
tag=${1:-r2k}
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
echo "=== parity ws (512)"; ASD_MC_NT=512 ASD_DEBUG=1 timeout 900 python -m pytest tests/test_gpu_mc_parity.py -m gpu -q -k lattice -W ignore 2>&1 | tail -6
echo "=== parity 256"; timeout 900 python -m pytest tests/test_gpu_mc_parity.py -m gpu -q -W ignore 2>&1 | tail -3
echo "=== mcbench 256"; timeout 600 python scripts/mcbench.py 128 128 128
echo "=== mcbench ws 512"; ASD_MC_NT=512 ASD_DEBUG=1 timeout 600 python scripts/mcbench.py 128 128 128
echo "=== mcbench 256x256x128 256"; timeout 600 python scripts/mcbench.py 256 256 128
echo "=== mcbench 256x256x128 ws"; ASD_MC_NT=512 timeout 600 python scripts/mcbench.py 256 256 128
