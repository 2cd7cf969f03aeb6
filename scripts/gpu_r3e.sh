# union walk: pad-2 rows, two entries per iteration (walk2.so) against four per iteration + closing pair (default build)
mkdir -p gpurun_out
exec > gpurun_out/r3e.log 2>&1
AB_REPS=2 python scripts/abbench.py uppasd_b200/libuppasd_b200.so build_var/walk2.so
AB_TEMP=0 AB_REPS=1 python scripts/abbench.py uppasd_b200/libuppasd_b200.so build_var/walk2.so
AB_SOLVER=5 AB_REPS=1 python scripts/abbench.py uppasd_b200/libuppasd_b200.so build_var/walk2.so
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
