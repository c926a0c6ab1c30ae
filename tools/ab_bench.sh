#!/bin/bash
# A/B of an environment knob through bench.py: tools/ab_bench.sh VAR A B [repeats]; prints value, ms/step and selected op times.
VAR=$1; A=$2; B=$3; N=${4:-2}
for i in $(seq $N); do for v in $A $B; do
  env $VAR=$v L3AC_BENCH_C_ABI=0 python bench.py --no-cpu-baseline --steps 12 --warmup 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$VAR=$v', round(d['value']), round(d['ms_per_step'],3), {k: d['op_ms'][k] for k in ('dwconv7_ln','convunit_mlp_tc','gemm_tc')}, d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done; done
