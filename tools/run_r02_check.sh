# round-2 one-GPU regression + measurement pass
mkdir -p gpurun_out
T=${1:-r02_run1}
{
  timeout 1700 python -m pytest tests -m gpu -q --maxfail=10 --durations=12 2>&1 | tail -60
  timeout 600 python __graft_entry__.py smoke 2>&1 | tail -4
  timeout 900 python bench.py 2>gpurun_out/${T}_bench_stderr.log | tee gpurun_out/${T}_bench.json
  tail -5 gpurun_out/${T}_bench_stderr.log
  timeout 600 python tools/bringup.py perf 2>&1 | grep PERF
  timeout 600 python tools/gemm_ab.py 2>&1 | tee gpurun_out/${T}_gemm_ab.log
} > gpurun_out/${T}.log 2>&1
cat gpurun_out/${T}.log
