# round-2 one-GPU pass: new-kernel tests, bench (bf16 headline + fp8 key), VAE-only bench, UMT5 timing, GEMM A/B, ncu
mkdir -p gpurun_out
T=${1:-r02_run3}
NCU="ncu --set full --clock-control none --import-source on --launch-skip 2 --launch-count 1"
{
  echo "== tests"; timeout 400 python -m pytest tests/test_t5_gpu.py tests/test_gemm_flex_gpu.py tests/test_fp8_gpu.py tests/test_server_loop_gpu.py "tests/test_dit_gpu.py::test_kv_roll_is_a_bit_exact_overlapping_memmove" tests/test_dit_gpu.py::test_eviction_branch -m gpu -q --maxfail=12 2>&1 | tail -40
  echo "== bench"; timeout 600 python bench.py 2>gpurun_out/${T}_bench_stderr.log | tee gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench_stderr.log
  echo "== vae bench"; timeout 200 python bench.py --workload vae_decode --steps 6 2>gpurun_out/${T}_vae_stderr.log | tee gpurun_out/${T}_bench_vae.json; tail -3 gpurun_out/${T}_vae_stderr.log
  echo "== t5"; timeout 200 python tools/time_t5.py 2>&1 | tail -12 | tee gpurun_out/${T}_umt5_timing.log
  echo "== gemm ab"; timeout 200 python tools/gemm_ab.py 2>&1 | tee gpurun_out/${T}_gemm_ab.log
  echo "== ncu"
  timeout 150 $NCU -k regex:gemm_flex -o gpurun_out/prof_r02_gemm_flex -f python tools/profile_kernels.py flex 2>&1 | tail -1
  timeout 150 $NCU -k regex:gemm_fp8 -o gpurun_out/prof_r02_gemm_fp8 -f python tools/profile_kernels.py fp8 2>&1 | tail -1
  timeout 150 $NCU -k regex:t5_attn -o gpurun_out/prof_r02_t5attn -f python tools/profile_kernels.py t5attn 2>&1 | tail -1
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 6800 -c 2800 --csv --log-file gpurun_out/launches_r02.csv python bench.py --layers 8 --steps 1 --warmup 3 --no-cpu-baseline --no-egress --no-fp8 > gpurun_out/bench_under_ncu_r02.json 2> gpurun_out/bench_under_ncu_r02.err
  tail -2 gpurun_out/bench_under_ncu_r02.err
} > gpurun_out/${T}.log 2>&1
tail -150 gpurun_out/${T}.log
