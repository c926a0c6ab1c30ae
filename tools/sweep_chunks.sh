#!/bin/bash
# Sweeps the micro-batch size and the number of streams of the 64 x 10 s bench step (tuning probe).
for cs in ${CHUNKS:-110 160 220 330}; do for st in ${STREAMS:-3 4 6}; do
  L3AC_CHUNK_SECONDS=$cs L3AC_STREAMS=$st timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunk_s=$cs streams=$st', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']))"
done; done
