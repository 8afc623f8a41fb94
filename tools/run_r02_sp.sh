# multi-GPU pass: P2P exchange check + single-stream bench (gpurun --gpus N -- bash tools/run_r02_sp.sh N tag)
N=${1:-2}; T=${2:-r02_sp$N}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
{
  nvidia-smi topo -m 2>&1 | head -14
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/check_sp.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -40
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 2>gpurun_out/${T}_bench_stderr.log | tee gpurun_out/${T}_bench.json
  tail -15 gpurun_out/${T}_bench_stderr.log
  if [ "${3:-}" = "ab" ]; then
    KR_GEMM_SK=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --no-secondary --no-cpu-baseline 2>gpurun_out/${T}_bench_sk_stderr.log | tee gpurun_out/${T}_bench_sk.json
    tail -5 gpurun_out/${T}_bench_sk_stderr.log
  fi
} > gpurun_out/${T}.log 2>&1
cat gpurun_out/${T}.log
