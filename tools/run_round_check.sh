# one-GPU regression + measurement pass (writes gpurun_out/round_check.log and bench JSON)
mkdir -p gpurun_out
{
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  timeout 600 python __graft_entry__.py smoke 2>&1 | tail -4
  timeout 900 python bench.py 2>gpurun_out/bench_stderr.log | tee gpurun_out/bench_latest.json
  tail -5 gpurun_out/bench_stderr.log
} > gpurun_out/round_check.log 2>&1
cat gpurun_out/round_check.log
