timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -x -q 2>&1 | tail -2
for lib in "" ab_libs/lib_head.so; do
  echo "lib=$lib"
  DMME_LIB_PATH=$lib python tools/prof_step.py 2>&1 | grep "qkv\|attn\|graph replay"
done
for rep in 1 2; do for lib in "" ab_libs/lib_head.so; do DMME_LIB_PATH=$lib python bench.py --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lib=$lib', d['ms_per_step'], d['value'])"; done; done
