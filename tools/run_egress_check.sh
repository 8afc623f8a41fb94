mkdir -p gpurun_out
{
  timeout 600 python -m pytest tests/test_egress_gpu.py -m gpu -x -q 2>&1 | tail -3
  timeout 600 python bench.py --no-cpu-baseline 2>gpurun_out/bench_stderr.log | tee gpurun_out/bench_egress.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value'],3),'e2e',round(d['e2e']['value'],3),d['e2e']['d2h_bytes_per_step'],'egress',d['egress_rgb8'])"
  tail -3 gpurun_out/bench_stderr.log
} > gpurun_out/egress_check.log 2>&1
cat gpurun_out/egress_check.log
