mkdir -p gpurun_out
exec > gpurun_out/kbench1.log 2>&1
for minb in 2 3 4; do
  for chunk in 8; do
    ASD_MINB=$minb ASD_CHUNK=$chunk python uppasd_b200/build.py || exit 1
    for pf in 0 444 1200; do
      ASD_MINB=$minb ASD_CHUNK=$chunk ASD_VARIANT=3 ASD_PF=$pf python scripts/kbench.py
    done
  done
done
ASD_MINB=2 ASD_CHUNK=4 python uppasd_b200/build.py && ASD_MINB=2 ASD_CHUNK=4 ASD_VARIANT=3 ASD_PF=444 python scripts/kbench.py
ASD_MINB=2 ASD_CHUNK=12 python uppasd_b200/build.py && ASD_MINB=2 ASD_CHUNK=12 ASD_VARIANT=3 ASD_PF=444 python scripts/kbench.py
ASD_MINB=2 ASD_CHUNK=16 python uppasd_b200/build.py && ASD_MINB=2 ASD_CHUNK=16 ASD_VARIANT=3 ASD_PF=444 python scripts/kbench.py
ASD_MINB=3 ASD_CHUNK=8 python uppasd_b200/build.py
ASD_VARIANT=0 python scripts/kbench.py
echo "=== pytest"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
