timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -x -q 2>&1 | tail -2
for mode in 1 17; do
  echo "halo_mode=$mode (bit4: res chunks last)"
  DMME_HALO_MODE=$mode python tools/prof_conv.py --only halo 2>&1 | head -9
  DMME_HALO_MODE=$mode python tools/prof_conv.py --only halo --gn 1 2>&1 | head -9
done
for rep in 1 2; do for mode in 1 17; do DMME_HALO_MODE=$mode python bench.py --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mode', $mode, d['ms_per_step'], d['value'])"; done; done
