# LEAN loop with the anisotropy test: headline A/B against the previous build (walk2.so ~ 0.418) + the new parity test + suite
mkdir -p gpurun_out
exec > gpurun_out/r3f.log 2>&1
AB_REPS=2 python scripts/abbench.py uppasd_b200/libuppasd_b200.so build_var/walk2.so
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
