# round 2: MC block sweep, run form: chain parity + micro-benchmark (A/B: tickets on / off, 512 / 256 threads)
tag=${1:-r2c}
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
echo "=== parity"; ASD_DEBUG=1 timeout 900 python -m pytest tests/test_gpu_mc_parity.py tests/test_gpu_mc_configs.py -m gpu -q 2>&1 | tail -8
echo "=== parity (512 threads)"; ASD_MC_NT=512 timeout 900 python -m pytest tests/test_gpu_mc_parity.py -m gpu -q -k lattice 2>&1 | tail -4
echo "=== parity (no tickets)"; ASD_MC_TICKET=0 timeout 900 python -m pytest tests/test_gpu_mc_parity.py -m gpu -q -k lattice 2>&1 | tail -4
echo "=== mcbench tickets"; ASD_DEBUG=1 timeout 600 python scripts/mcbench.py 128 128 128
echo "=== mcbench tickets 512"; ASD_MC_NT=512 timeout 600 python scripts/mcbench.py 128 128 128
echo "=== mcbench no tickets"; ASD_MC_TICKET=0 timeout 600 python scripts/mcbench.py 128 128 128
echo "=== mcbench 256x256x128"; timeout 600 python scripts/mcbench.py 256 256 128
echo "=== mcbench 64"; ASD_DEBUG=1 timeout 600 python scripts/mcbench.py 64 64 64
echo "=== phases"; export ASD_LIB=$PWD/build_var/libprof.so; timeout 300 python scripts/mcprof.py 128 128 128; ASD_MC_NT=512 timeout 300 python scripts/mcprof.py 128 128 128
