mkdir -p gpurun_out
{
  for mode in 1 4 1 4; do
    echo "== KR_GEMM2=$mode"
    KR_GEMM2=$mode timeout 600 python bench.py --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],3),'fps', round(d['ms_per_step'],1),'ms', d['clocks']['sm_mhz'],'MHz', 'gemm frac', round(d['roofline']['frac'],4), {k:round(v['frac'],3) for k,v in d['roofline']['by_kernel'].items()})"
  done
} > gpurun_out/gemm_hybrid_ab.log 2>&1
cat gpurun_out/gemm_hybrid_ab.log
