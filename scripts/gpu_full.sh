mkdir -p gpurun_out
exec > gpurun_out/fullham.log 2>&1
echo "=== bench do_reduced N"; timeout 600 python bench.py --full-ham --steps 200 --warmup 5 --no-cpu
echo "=== pytest cabi + parity subset"; timeout 600 python -m pytest tests/test_cabi.py tests/test_gpu_parity.py -m gpu -q -k "c_example or fixed_moment or per_site" 2>&1 | tail -4
