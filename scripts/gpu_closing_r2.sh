mkdir -p gpurun_out
exec > gpurun_out/closing_r2.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
timeout 400 python bench.py 2>&1 | tail -1 | tee gpurun_out/closing_r2_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/closing_r2_ref.json
