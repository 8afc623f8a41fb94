#!/bin/bash
# last GPU pass of round 2 (3 GPU-minutes left): the device JPEG encoder vs Pillow, then its timing
mkdir -p gpurun_out
timeout 110 python -m pytest tests/test_zz_jpeg_gpu.py -x -q > gpurun_out/jpeg_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/jpeg_tests.log
tail -5 gpurun_out/jpeg_tests.log
timeout 50 python tools/jpeg_time.py > gpurun_out/jpeg_time.log 2>&1
echo "time exit $?" >> gpurun_out/jpeg_time.log
tail -3 gpurun_out/jpeg_time.log
