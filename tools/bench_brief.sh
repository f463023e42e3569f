#!/bin/bash
# usage: tools/bench_brief.sh <lib.so> [bench args]  -- one-line summary of a bench.py run with a given library build
lib=$1; shift
ZKMSM_DEV=1 ZKMSM_LIB=$lib timeout 300 python bench.py --no-cpu-baseline --no-extras "$@" 2>&1 | tail -1 | python3 -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$lib'.split('/')[-1], 'value', round(d['value']/1e6,1), 'ms', round(d['ms_per_step'],3), 'lat', round(d['impl_config']['single_msm_latency_ms'],3), 'e2e', round(d['e2e']['value']/1e6,1), round(d['e2e']['ms_per_step'],3), 'pre', round(d.get('precomputed_tables',{}).get('ms_per_step',0),3))"
