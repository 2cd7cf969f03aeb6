mkdir -p gpurun_out
exec > gpurun_out/var1.log 2>&1
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scripts/kbench.py
ASD_PRELOAD=0 python scripts/kbench.py
ASD_KEYWRAP=0 python scripts/kbench.py
ASD_KEYWRAP=0 ASD_PRELOAD=0 python scripts/kbench.py
