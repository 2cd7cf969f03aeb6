# round 2: compute-sanitizer on the new kernels (MC block sweep run form with tickets, generic block kernel, sd_run, time field, async legacy loop)
tag=${1:-r2m}
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
SEL='warp_specialised or (lattice_layouts and bccfe and 1024) or (lattice_layouts and kagome and 256) or block_sweep_observables'
echo "=== memcheck MC"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_mc_parity.py -m gpu -q -x -W ignore -k "$SEL" 2>&1 | tail -8
echo "=== racecheck MC"; timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_mc_parity.py -m gpu -q -x -W ignore -k "warp_specialised or (lattice_layouts and bccfe and 1024)" 2>&1 | tail -8
echo "=== memcheck LLG extras"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_alloy.py -m gpu -q -x -W ignore -k "time_dependent or asynchronously or pyasd or (alloy_tables and ncell0)" 2>&1 | tail -8
echo "=== synccheck MC"; timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_mc_parity.py -m gpu -q -x -W ignore -k "warp_specialised or (lattice_layouts and bccfe and 1024)" 2>&1 | tail -8
