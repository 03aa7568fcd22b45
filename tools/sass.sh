#!/bin/bash
# Build container: compile pnb_eval.cu alone to a cubin with the given -D flags and dump the SASS of the
# pruning kernel (FUSED=false, RESCALE=false) to /tmp/prune.sass; print register use and an opcode histogram.
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -cubin "$@" -Xptxas -v \
  phynetpy_b200/csrc/pnb_eval.cu -o /tmp/pnb_eval.cubin 2>&1 | grep -A2 "pnb_prune_kernel_tILb0ELb0" | grep -E "registers|spill"
cuobjdump -sass -fun '_ZN3pnb18pnb_prune_kernel_tILb0ELb0EEEvNS_11PruneLaunchENS_11ModelParamsE' /tmp/pnb_eval.cubin > /tmp/prune.sass
grep -cE "^\s+/\*[0-9a-f]{4}\*/" /tmp/prune.sass
for op in BSSY BSYNC "BRA" "BRX" "LDS" "DFMA" "DMUL" "LOP3" "ISETP" "UISETP" "VOTE" "R2UR" "WARPSYNC" "BREAK"; do
  printf "%s=%s " $op $(grep -cE "\b$op" /tmp/prune.sass)
done; echo
