# gpurun --gpus N -- bash tools/run_r02_sp_bench_only.sh N tag  : single-stream bench only (short box time)
N=${1:-8}; T=${2:-r02_sp$N}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 --no-secondary --no-cpu-baseline 2>gpurun_out/${T}_bench_stderr.log | tee gpurun_out/${T}_bench.json
grep -v "^W0\|^\*\*\*\|Setting OMP" gpurun_out/${T}_bench_stderr.log | tail -8
