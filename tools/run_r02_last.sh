# last round-2 GPU pass (one B200, short): fixed plan test, DRAM traffic of a VAE decode block, r02 captures of the two
# kernels the step spends most time in (unchanged since round 1; fresh numbers on this build)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --launch-skip 2 --launch-count 1"
{
  timeout 120 python -m pytest tests/test_gemm_streamk_gpu.py tests/test_gemm_flex_gpu.py -m gpu -q 2>&1 | tail -3
  timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/vae_dram_r02.csv python bench.py --workload vae_decode --steps 2 --warmup 3 > gpurun_out/vae_under_ncu.json 2> gpurun_out/vae_under_ncu.err; tail -2 gpurun_out/vae_under_ncu.err
  timeout 120 $NCU -k regex:attn_fwd -o gpurun_out/prof_r02_attn -f python tools/profile_kernels.py attn 2>&1 | tail -1
  timeout 120 $NCU -k regex:gemm_tn -o gpurun_out/prof_r02_gemm1_ffn2 -f python tools/profile_kernels.py gemm1 2>&1 | tail -1
} > gpurun_out/r02_last.log 2>&1
cat gpurun_out/r02_last.log
