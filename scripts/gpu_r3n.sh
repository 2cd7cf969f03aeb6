# compute-sanitizer over the moment-plane run kernels: memcheck on the run-kernel / slab tests, racecheck + synccheck on a subset
tag=r3n
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
echo "=== memcheck: tests/test_gpu_lattice.py + slab planes"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_slab.py -m gpu -x -q -k "run_kernel or planes or fcc or headline_launch_geometry_64 or fixed_moment" 2>&1 | tail -6
echo "=== racecheck: moment planes + anisotropy + fcc"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_lattice.py -m gpu -x -q -k "moment_planes or anisotropy_and_field" 2>&1 | tail -6
echo "=== synccheck"
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_lattice.py -m gpu -x -q -k "moment_planes" 2>&1 | tail -6
echo "=== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
