# one-GPU regression + measurement pass (writes gpurun_out/round_check.log and bench JSON)
mkdir -p gpurun_out
{
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  timeout 600 python __graft_entry__.py smoke 2>&1 | tail -4
  timeout 900 python bench.py 2>gpurun_out/bench_stderr.log | tee gpurun_out/bench_latest.json
  tail -5 gpurun_out/bench_stderr.log
  timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 2 --launch-count 1 -k regex:conv_halo -o gpurun_out/prof_r01c_conv_halo -f python tools/profile_kernels.py conv 2>&1 | tail -2
} > gpurun_out/round_check.log 2>&1
cat gpurun_out/round_check.log
