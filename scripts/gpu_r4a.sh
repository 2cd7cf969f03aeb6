# lean integrator in the staged kernel with one coupling row per atom (do_reduced N): bench --full-ham, alloybench; suite
mkdir -p gpurun_out
exec > gpurun_out/r4a.log 2>&1
run() { python bench.py --full-ham --steps 20 --warmup 3 --no-cpu --no-secondary 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('full-ham', 'ms', d['ms_per_step'], 'value', d['value'], d['roofline']['kernel'])
"; }
echo "lean"; run
echo "general"; ASD_LEAN=0 run
echo "=== alloy"; python scripts/alloybench.py 2>&1 | tail -4
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -2
