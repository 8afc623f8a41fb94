# A/B of the attention kernels + GEMM perf on one B200 (writes gpurun_out/attn_ab.log)
mkdir -p gpurun_out
{
  echo "== db"
  timeout 300 python tools/bringup.py attnperf 2>&1 | grep PERF
  echo "== pingpong"
  KR_ATTN_KERNEL=pingpong timeout 300 python tools/bringup.py attnperf 2>&1 | grep PERF
  echo "== perf (gemm)"
  timeout 300 python tools/bringup.py perf 2>&1 | grep PERF
  KR_GEMM2=0 timeout 300 python tools/bringup.py perf 2>&1 | grep "PERF gemm" | sed 's/^/1cta /'
  timeout 300 python tools/bringup.py attn 2>&1 | grep -E "PASS|FAIL" | tail -3
  KR_ATTN_KERNEL=pingpong timeout 300 python tools/bringup.py attn 2>&1 | grep -E "PASS|FAIL" | tail -3
  timeout 300 python tools/bringup.py gemm 2>&1 | grep -E "PASS|FAIL" | tail -3
  timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
} > gpurun_out/attn_ab.log 2>&1
cat gpurun_out/attn_ab.log
