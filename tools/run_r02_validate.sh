# final one-GPU validation of the round-2 tree: full GPU suite, smoke, short bench
mkdir -p gpurun_out
{
  timeout 420 python -m pytest tests -m gpu -q --maxfail=10 --durations=6 2>&1 | tail -25
  timeout 120 python __graft_entry__.py smoke 2>&1 | tail -4
  timeout 300 python bench.py --steps 3 --no-cpu-baseline 2>gpurun_out/r02_validate_bench_stderr.log | tee gpurun_out/r02_validate_bench.json | cut -c1-400
  tail -3 gpurun_out/r02_validate_bench_stderr.log
} > gpurun_out/r02_validate.log 2>&1
cat gpurun_out/r02_validate.log
