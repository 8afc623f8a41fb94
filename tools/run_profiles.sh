# ncu captures of the hot kernels + launch list of one bench step (one B200; outputs under gpurun_out/)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --launch-skip 2 --launch-count 1"
{
  timeout 900 python -m pytest tests/test_attention_gpu.py -m gpu -x -q 2>&1 | tail -3
  timeout 600 $NCU -k regex:gemm2 -o gpurun_out/prof_r01b_gemm2_ffn1 -f python tools/profile_kernels.py gemm 2>&1 | tail -2
  timeout 600 $NCU -k regex:gemm_tn -o gpurun_out/prof_r01b_gemm1_ffn2 -f python tools/profile_kernels.py gemm1 2>&1 | tail -2
  timeout 600 $NCU -k regex:attn -o gpurun_out/prof_r01b_attn -f python tools/profile_kernels.py attn 2>&1 | tail -2
  timeout 600 $NCU -k regex:conv_igemm -o gpurun_out/prof_r01b_conv -f python tools/profile_kernels.py conv 2>&1 | tail -2
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 1 --warmup 1 > gpurun_out/bench_under_ncu_b.json 2> gpurun_out/bench_under_ncu_b.err
  tail -2 gpurun_out/bench_under_ncu_b.err
} > gpurun_out/profiles_run.log 2>&1
cat gpurun_out/profiles_run.log
