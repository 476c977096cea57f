#!/bin/bash
# A/B of differently compiled builds of the same sources (development only): scripts/dev_fused.py with the
# in-tree library and with every auromat_b200/csrc/_variants/*.so (see auromat_b200/_lib.py AMT_LIB).
#   gpurun --timeout 900 -- 'bash scripts/gpu_ab.sh tag'
set -u
tag=${1:-ab}
out=gpurun_out
mkdir -p $out
timeout 300 python scripts/dev_fused.py > $out/ab_${tag}_default.log 2>&1; echo "default rc=$?"
AMT_SEQ_TRACE=1 timeout 300 python scripts/seq_trace.py 60 > $out/ab_${tag}_default.trace 2>&1
for v in auromat_b200/csrc/_variants/*.so; do
  [ -f "$v" ] || continue
  AMT_LIB=$PWD/$v timeout 300 python scripts/dev_fused.py > $out/ab_${tag}_$(basename $v .so).log 2>&1; echo "$v rc=$?"
  AMT_LIB=$PWD/$v AMT_SEQ_TRACE=1 timeout 300 python scripts/seq_trace.py 60 > $out/ab_${tag}_$(basename $v .so).trace 2>&1
done
for f in $out/ab_${tag}_*.log; do echo $f; grep -h "timings" $f | grep -o "'fused planes+mag+bin': [0-9.]*\|'fused bin only': [0-9.]*\|'fused planes+mag': [0-9.]*" | tr '\n' ' '; echo; grep -ci "mismatch\|differ\|FAIL" $f; grep -E "^==|^median" ${f%.log}.trace; done
