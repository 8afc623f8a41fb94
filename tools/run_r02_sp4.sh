# gpurun --gpus 4 -- bash tools/run_r02_sp4.sh : P2P exchange at 4 ranks + single-stream bench, tight limits
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
{
  timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 tools/check_sp.py 2>&1 | grep "SP CHECK\|rel_l2" | tail -12
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 3 --warmup 3 --no-secondary --no-cpu-baseline --watchdog-s 150 2>gpurun_out/r02_sp4_bench_stderr.log | tee gpurun_out/r02_sp4_bench.json
  grep -v "^W0\|^\*\*\*\|Setting OMP" gpurun_out/r02_sp4_bench_stderr.log | tail -6
} > gpurun_out/r02_sp4.log 2>&1
cat gpurun_out/r02_sp4.log
