mkdir -p gpurun_out
exec > gpurun_out/pdl.log 2>&1
for v in 1 0 1 0; do
  echo "=== ASD_PDL=$v"; ASD_PDL=$v python bench.py --steps 500 --warmup 10 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
echo "=== ASD_PDL=1 T=0"; ASD_PDL=1 python bench.py --steps 500 --warmup 10 --no-cpu --temp 0 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
echo "=== pytest (PDL on)"; timeout 900 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_slab.py tests/test_gpu_skyrmion.py -m gpu -q 2>&1 | tail -4
echo "=== smoke"; python __graft_entry__.py smoke 2>&1 | tail -2
