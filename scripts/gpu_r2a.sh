# round 2, first GPU call: the new parity tests (Monte Carlo chain parity incl. the block sweep, torque, headline geometry),
# then the MC micro-benchmark on 64^3 / 128^3 (block sweep vs colour-major)
tag=${1:-r2a}
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
echo "=== new tests"; ASD_DEBUG=1 timeout 1500 python -m pytest tests/test_gpu_mc_parity.py tests/test_gpu_parity.py::test_spin_transfer_torque_field tests/test_gpu_lattice.py -m gpu -q -x 2>&1 | tail -40
echo "=== mcbench block"; ASD_DEBUG=1 timeout 600 python scripts/mcbench.py 64 64 64; ASD_DEBUG=1 timeout 600 python scripts/mcbench.py 128 128 128
echo "=== mcbench block ts 256"; ASD_MC_TS=256 timeout 600 python scripts/mcbench.py 128 128 128
echo "=== mcbench colour-major"; ASD_MC_BLOCK=0 timeout 600 python scripts/mcbench.py 128 128 128
