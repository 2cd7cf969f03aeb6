# round 2: compute-sanitizer memcheck over the whole GPU suite (minus the long statistical scans)
tag=${1:-r2r}
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
timeout 3000 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -W ignore -x -k "not temperature_scan and not block_sweep_observables and not thermal and not golden_on_gpu and not run_directory and not run_directories and not cumulant_goldens" 2>&1 | tail -12
