mkdir -p gpurun_out
exec > gpurun_out/unroll.log 2>&1
for v in default u2 u8 default; do
  if [ $v = default ]; then unset ASD_LIB; else export ASD_LIB=/root/repo/build_var/lib_$v.so; fi
  echo "=== $v"; python bench.py --steps 400 --warmup 10 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['clocks'])"
done
