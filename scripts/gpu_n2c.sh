# 2-GPU sanity run at the end of the round: slab tests across devices, both bench arms under torchrun
tag=${1:-n2c}
N=${2:-2}
mkdir -p gpurun_out
exec > gpurun_out/$tag.log 2>&1
echo "=== pytest slab"; timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "=== ensemble x$N"; timeout 900 $TR bench.py --gpus $N --steps 300 --warmup 10
echo "=== reference arm x$N"; timeout 900 $TR bench.py --impl reference --gpus $N --steps 3 --warmup 1
echo "=== slab x$N 128^3"; timeout 900 $TR bench.py --gpus $N --steps 300 --warmup 10 --decomp slab --no-cpu
