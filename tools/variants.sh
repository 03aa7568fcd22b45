#!/bin/bash
# A/B harness (GPU box): rebuild the library with different tuning knobs and run the cfg2 bench for each.
mkdir -p gpurun_out
for v in "$@"; do
  export PNB_NVCC_DEFS="$(echo $v | tr ',' ' ')"
  python -m phynetpy_b200.build --force > /dev/null 2>gpurun_out/build_$$.log || { echo "BUILD FAILED $v"; tail -5 gpurun_out/build_$$.log; continue; }
  echo "=== $v : $(grep -A2 prune_kernel gpurun_out/build_$$.log | grep -oE 'Used [0-9]+ registers' | head -1) $(grep -A2 prune_kernel gpurun_out/build_$$.log | grep -oE '[0-9]+ bytes spill stores' | head -1)"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('   value=%.3e e2e=%.3e ms/step=%.4f kernel_ms=%.4f frac=%.3f logL=%.6f'%(d['value'],d['e2e']['value'],d['ms_per_step'],r['kernel_ms'],r['frac'],d['logL']))"
done
unset PNB_NVCC_DEFS
python -m phynetpy_b200.build --force > /dev/null 2>&1
