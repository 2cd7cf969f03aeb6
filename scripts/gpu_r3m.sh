# eight GPUs: the driver's command under torchrun (ensembles + secondary.slab 512x512x256 + secondary.mc / config3 / config4)
N=${1:-8}
mkdir -p gpurun_out
exec > gpurun_out/r3m_n$N.log 2>&1
nvidia-smi --query-gpu=index,name --format=csv | head -10
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
echo "=== driver command x$N"; timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 2>&1 | grep -E '^\{|rror' | tee gpurun_out/bench20_r3m_n$N.json | cut -c1-300
