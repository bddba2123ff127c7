timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for lib in "" ab_libs/lib_p2.so; do
  echo "lib=$lib"
  DMME_LIB_PATH=$lib python tools/prof_conv.py --only halo --gn 1 2>&1 | grep res
done
for rep in 1 2; do for lib in "" ab_libs/lib_p2.so; do DMME_LIB_PATH=$lib python bench.py --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lib=$lib', d['ms_per_step'], d['value'])"; done; done
