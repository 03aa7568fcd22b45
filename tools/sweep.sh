#!/bin/bash
# GPU box: run the cfg2 bench (and optionally another workload) once per prebuilt library variant.
#   tools/sweep.sh cfg2 base t64p4b8 ...      -> gpurun_out/sweep_<workload>.txt
wl=$1; shift
mkdir -p gpurun_out
out=gpurun_out/sweep_$wl.txt
: > $out
for v in "$@"; do
  lib=phynetpy_b200/variants/lib_$v.so
  [ -f $lib ] || { echo "$v: missing $lib" | tee -a $out; continue; }
  PNB_LIB=$PWD/$lib timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline ${SWEEP_ARGS} 2> gpurun_out/sweep_err_$v.log | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    r=d['roofline']; s=d.get('score_only') or {}
    print('$v value=%.4e e2e=%.4e ms/step=%.4f kernel_ms=%.4f score_ms=%s pm_ms=%.4f logL=%.9f' % (d['value'],d['e2e']['value'],d['ms_per_step'],r['kernel_ms'],('%.4f'%s['kernel_ms']) if s else 'NA',r['pmatrix_kernel_ms'],d['logL']))
except Exception as e:
    print('$v FAILED', e)
" | tee -a $out
done
