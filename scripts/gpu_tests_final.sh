mkdir -p gpurun_out
exec > gpurun_out/tests_final.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
