# closing run of the session: smoke, bench line, reference arm line on the final build
mkdir -p gpurun_out
exec > gpurun_out/r3w.log 2>&1
echo "=== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "=== bench 20"; python bench.py --steps 20 --warmup 3 | tee gpurun_out/bench20_r3w.json | cut -c1-200
echo "=== reference arm"; python bench.py --impl reference --steps 20 --warmup 3 | tee gpurun_out/ref20_r3w.json | cut -c1-300
