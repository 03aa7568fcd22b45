#!/bin/bash
# GPU box: cfg2 bench for (variant, env) pairs: tools/sweep_env.sh cfg2 "variant ENV=VAL ..." ...
wl=$1; shift
mkdir -p gpurun_out
out=gpurun_out/sweep_env_$wl.txt
: > $out
for spec in "$@"; do
  set -- $spec
  v=$1; shift
  lib=phynetpy_b200/variants/lib_$v.so
  [ -f $lib ] || { echo "$v: missing $lib" | tee -a $out; continue; }
  env PNB_LIB=$PWD/$lib "$@" timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline ${SWEEP_ARGS} 2> gpurun_out/sweep_err.log | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    r=d['roofline']; s=d.get('score_only') or {}
    print('%-34s value=%.4e e2e=%.4e ms/step=%.4f kernel_ms=%.4f score_ms=%s logL=%.9f' % ('$spec',d['value'],d['e2e']['value'],d['ms_per_step'],r['kernel_ms'],('%.4f'%s['kernel_ms']) if s else 'NA',d['logL']))
except Exception as e:
    print('$spec FAILED', e)
" | tee -a $out
done
