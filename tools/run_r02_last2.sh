#!/bin/bash
# final GPU pass of round 2 (2.2 GPU-minutes left): conv vs cuDNN + FP8 GEMM vs cuBLASLt A/B, then the ncu launch
# list of the device JPEG encoder's four passes
mkdir -p gpurun_out
timeout 50 python tools/conv_fp8_ab.py > gpurun_out/conv_fp8_ab.log 2>&1
echo "ab exit $?" >> gpurun_out/conv_fp8_ab.log
tail -8 gpurun_out/conv_fp8_ab.log
timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jpeg_ -c 32 --csv \
  --log-file gpurun_out/jpeg_launches.csv python tools/jpeg_time.py > gpurun_out/jpeg_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/jpeg_ncu.log
tail -3 gpurun_out/jpeg_ncu.log
