tag=r3o
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
echo "=== memcheck: run-kernel tests (tests/test_gpu_lattice.py)"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_lattice.py -m gpu -x -q -k "run_kernel or planes or fcc or headline_launch_geometry_64 or fixed_moment" 2>&1 | tail -6
echo "=== memcheck: slab on moment planes (in-process ring; the sanitizer serialises kernels of different streams)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_slab.py -m gpu -x -q -k "planes" 2>&1 | grep -E "Error|error|passed|failed|SUMMARY|timed" | head -12
