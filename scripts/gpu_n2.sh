tag=${1:-n2}
N=${2:-2}
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
nvidia-smi --query-gpu=index,name --format=csv
nvidia-smi topo -m | head -12
echo "=== pytest slab"; timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "=== ensemble x$N"; timeout 900 $TR bench.py --gpus $N --steps 300 --warmup 10
echo "=== slab x$N 128^3"; timeout 900 $TR bench.py --gpus $N --steps 300 --warmup 10 --decomp slab
echo "=== slab x$N 256x256x256"; timeout 900 $TR bench.py --gpus $N --steps 100 --warmup 5 --decomp slab --ncell 256 256 256
echo "=== N=1 256^3 for reference"; timeout 900 python bench.py --steps 100 --warmup 5 --ncell 256 256 256 --no-cpu
