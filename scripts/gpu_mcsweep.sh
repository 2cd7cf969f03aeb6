mkdir -p gpurun_out
exec > gpurun_out/mcsweep.log 2>&1
for mb in 2 3 4; do for ch in 4 8; do
  ASD_MC_MINB=$mb ASD_MC_CHUNK=$ch python uppasd_b200/build.py > /dev/null || exit 1
  echo "--- MINB $mb CHUNK $ch"; python scripts/mcbench.py 128 128 128 | grep "^MC [MH]"
done; done
python uppasd_b200/build.py > /dev/null
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
