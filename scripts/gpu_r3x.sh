# union walk with split accumulators for classes of one (split1) / up to two (split2) runs, against the default build
mkdir -p gpurun_out
exec > gpurun_out/r3x.log 2>&1
AB_REPS=2 python scripts/abbench.py uppasd_b200/libuppasd_b200.so build_var/split1.so build_var/split2.so
AB_TEMP=0 AB_REPS=1 python scripts/abbench.py uppasd_b200/libuppasd_b200.so build_var/split1.so build_var/split2.so
echo "=== parity (split2)"; ASD_LIB=build_var/split2.so timeout 900 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -3
