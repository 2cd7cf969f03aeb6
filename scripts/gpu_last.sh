tag=${1:-last}
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "=== smoke"; python __graft_entry__.py smoke 2>&1 | tail -4
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["kernel"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])'
echo "=== bench (default)"; python bench.py --steps 1000 --warmup 10 | tee gpurun_out/bench_$tag.json | python -c "$P"
echo "=== staged kernel ASD_RUNS=0 PDL=1"; ASD_RUNS=0 python bench.py --steps 400 --warmup 10 --no-cpu | python -c "$P"
echo "=== staged kernel ASD_RUNS=0 PDL=0"; ASD_RUNS=0 ASD_PDL=0 python bench.py --steps 400 --warmup 10 --no-cpu | python -c "$P"
echo "=== depondt"; python bench.py --steps 400 --warmup 10 --solver 5 --no-cpu | python -c "$P"
echo "=== reference arm"; python bench.py --impl reference --steps 5 --warmup 1 | cut -c1-400
