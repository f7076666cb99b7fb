#!/bin/bash
# Developer tool (GPU box): one default bench run without the side workloads, printed in brief.
python bench.py --no-also --no-cpu-baseline --e2e-steps 1 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print(round(d['ms_per_step'],4), d['steps'], {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()}, 'sum', round(sum(v['ms_per_step'] for v in d['kernels'].values()),4))
"
