# PDL experiment (library built with EXTRA=-DMVLDM_ENABLE_PDL): plain stream order vs programmatic dependent launch with the
# dependents released late / at entry, with and without the weight L2 prefetch ahead of griddepcontrol.wait
B="python bench.py --steps 25 --no-cpu-baseline --no-config4 --view-sharded-views 0"
for cfg in "0 7" "1 7" "1 23" "1 31" "1 15"; do
  set -- $cfg
  MVLDM_PDL=$1 MVLDM_GEMM_OPT=$2 $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('PDL=$1 GEMM_OPT=$2: %.3f ms/step  %.1f steps/s  e2e %.1f' % (d['ms_per_step'], d['value'], d['e2e']['value']))"
done
MVLDM_PDL=1 MVLDM_GEMM_OPT=31 python -m pytest tests/test_gpu_forward.py -q -x -k "v8_golden or v4_per_layer or trajectory" 2>&1 | tail -3
