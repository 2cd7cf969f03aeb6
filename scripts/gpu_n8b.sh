mkdir -p gpurun_out
exec > gpurun_out/n8b.log 2>&1
nvidia-smi --query-gpu=index,name --format=csv | head -12
echo "=== pytest slab (8 devices)"; timeout 600 python -m pytest tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -3
run() { N=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" 2>&1 | grep -E '^\{|Error|error' ; }
echo "=== ensemble x8"; run 8 --steps 300 --warmup 10
echo "=== slab x8 512x512x256"; run 8 --steps 60 --warmup 5 --decomp slab --ncell 512 512 256
echo "=== slab x8 256^3"; run 8 --steps 100 --warmup 5 --decomp slab --ncell 256 256 256
echo "=== slab x4 512x512x256"; run 4 --steps 40 --warmup 5 --decomp slab --ncell 512 512 256
