mkdir -p gpurun_out
exec > gpurun_out/staged1.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
echo "=== pytest (default build)"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== kbench"
python scripts/kbench.py
ASD_STAGED=0 python scripts/kbench.py
ASD_PF=0 python scripts/kbench.py
for mb in 2 4; do
  ASD_MINB_STAGED=$mb python uppasd_b200/build.py || exit 1
  ASD_MINB_STAGED=$mb python scripts/kbench.py
done
for npf in 2 7; do
  ASD_MINB_STAGED=4 ASD_NPF=$npf python uppasd_b200/build.py || exit 1
  ASD_MINB_STAGED=4 ASD_NPF=$npf python scripts/kbench.py
done
python uppasd_b200/build.py
