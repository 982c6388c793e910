#!/bin/bash
# usage: scripts/exp.sh TAG WORKLOAD [ENV=VAL ...]   -- one bench.py run of one workload, summary line only
tag=$1; wl=$2; shift 2
env "$@" python bench.py --workload $wl --configs none --no-e2e --no-cpu --steps 5 --warmup 3 > gpurun_out/exp_$tag.log 2> gpurun_out/exp_$tag.err
rc=$?
python - "$tag" $rc <<'PY'
import json,sys
tag,rc=sys.argv[1],sys.argv[2]
try:
    l=json.loads(open('gpurun_out/exp_%s.log'%tag).read().strip().splitlines()[-1])
    p=l.get('parity',{})
    print(tag, 'rc',rc, 'step %.3f'%l['ms_per_step'], {k:round(v,3) for k,v in l['phase_ms_per_step'].items()}, 'err',p.get('max_rel_err'),'draws',p.get('draws_bit_exact'),'ok',p.get('ok'))
except Exception as e:
    print(tag,'rc',rc,'FAILED',e); print(open('gpurun_out/exp_%s.err'%tag).read()[-800:])
PY
