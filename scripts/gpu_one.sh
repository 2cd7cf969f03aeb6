mkdir -p gpurun_out
exec > gpurun_out/one.log 2>&1
python scripts/slab_selfring.py 2>&1 | tail -4
echo "=== memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/slab_selfring.py 2>&1 | tail -6
