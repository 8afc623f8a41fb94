mkdir -p gpurun_out
{
  echo "== halo (default)"
  timeout 900 python -m pytest tests/test_vae_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -4
  timeout 600 python tools/bringup_vae.py 2>&1 | grep call
  echo "== per-tap kernel (KR_CONV_HALO=0)"
  KR_CONV_HALO=0 timeout 600 python tools/bringup_vae.py 2>&1 | grep call
} > gpurun_out/vae_ab.log 2>&1
cat gpurun_out/vae_ab.log
