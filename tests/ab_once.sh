#!/bin/bash
# Developer tool (GPU box): like ab_bench.sh, one round.
for lib in "$@"; do
  PDABM_LIB=$PWD/$lib python bench.py --steps 30 --no-also --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$lib', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
"
done
