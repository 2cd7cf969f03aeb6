mkdir -p gpurun_out
exec > gpurun_out/one.log 2>&1
run() { python bench.py --ncell $1 $2 $3 --steps 20 --warmup 3 --no-cpu --no-secondary $4 $5 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 $2 $3 $4 $5', 'ms', d['ms_per_step'], 'frac', d['roofline']['step_frac_of_peak'], 'launches', d['gpu_launches'])
    elif 'rror' in l: print(l.strip()[:200])
"; }
echo "side stream, PDL"; run 512 512 32 --decomp slab
echo "side stream, no PDL"; ASD_SLAB_PDL=0 run 512 512 32 --decomp slab
echo "fused edge kernels, PDL on the interior launch"; ASD_SLAB_SIDE=0 run 512 512 32 --decomp slab
echo "fused edge kernels, no PDL"; ASD_SLAB_SIDE=0 ASD_SLAB_PDL=0 run 512 512 32 --decomp slab
echo "=== slab tests"; timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -2
ASD_SLAB_SIDE=0 timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -2
