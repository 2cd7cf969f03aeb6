# compute-sanitizer passes over the small-system kernels and the new observables (memcheck + racecheck)
tag=${1:-san}
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
SEL='resident or fixed_moment or observables_on_host or cluster_golden or energy_terms_parity'
echo "=== memcheck"; timeout 800 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mc_configs.py -m gpu -x -q -k "$SEL" 2>&1 | tail -15
echo "=== racecheck"; timeout 800 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mc_configs.py -m gpu -x -q -k "resident" 2>&1 | tail -15
