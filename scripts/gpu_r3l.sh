# two GPUs: slab tests across devices, the driver's command under torchrun (ensembles + secondary.slab), slab 256^3 alone
N=${1:-2}
mkdir -p gpurun_out
exec > gpurun_out/r3l_n$N.log 2>&1
nvidia-smi --query-gpu=index,name --format=csv | head -10
echo "=== pytest slab"; timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "=== driver command x$N"; timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 2>&1 | grep -E '^\{|rror' | tee gpurun_out/bench20_r3l_n$N.json | cut -c1-400
echo "=== slab x$N 256^3, 100 steps"; timeout 900 $TR bench.py --gpus $N --steps 100 --warmup 5 --decomp slab --ncell 256 256 256 --no-secondary 2>&1 | grep -E '^\{|rror' | cut -c1-600
