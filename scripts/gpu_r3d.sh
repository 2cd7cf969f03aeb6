# union-walk variants (mask classes padded to WALK_U entries, U entries per iteration): U = 1 (unpadded, compiler-unrolled), 2, 4
mkdir -p gpurun_out
exec > gpurun_out/r3d.log 2>&1
AB_REPS=2 python scripts/abbench.py build_var/walk1.so build_var/walk2.so build_var/walk4.so
AB_TEMP=0 AB_REPS=1 python scripts/abbench.py build_var/walk1.so build_var/walk2.so build_var/walk4.so
echo "=== parity of the padded rows (walk2, walk4): run-kernel tests"
ASD_LIB=build_var/walk2.so timeout 900 python -m pytest tests/test_gpu_lattice.py -m gpu -x -q 2>&1 | tail -3
ASD_LIB=build_var/walk4.so timeout 900 python -m pytest tests/test_gpu_lattice.py -m gpu -x -q 2>&1 | tail -3
