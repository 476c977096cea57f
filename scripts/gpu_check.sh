#!/bin/bash
# First GPU contact of a change: dev check of the fused kernels, the gpu test-suite, a short bench.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh tag'
set -u
tag=${1:-chk}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/smi_$tag.txt 2>&1
timeout 600 python scripts/dev_fused.py > $out/dev_fused_$tag.log 2>&1; echo "dev_fused rc=$?"
# A/B: other builds of the same sources (development only, see auromat_b200/_lib.py AMT_LIB)
for v in auromat_b200/csrc/_variants/*.so; do
  [ -f "$v" ] || continue
  AMT_LIB=$PWD/$v timeout 600 python scripts/dev_fused.py > $out/dev_fused_${tag}_$(basename $v .so).log 2>&1; echo "dev_fused $v rc=$?"
done
grep -h "timings" $out/dev_fused_${tag}*.log | cut -c1-400
timeout 1200 python -m pytest tests -q -m gpu -x --timeout 900 > $out/gputests_$tag.log 2>&1; echo "pytest rc=$?"
tail -5 $out/gputests_$tag.log
timeout 600 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
tail -c 3000 $out/bench_$tag.json
tail -5 $out/dev_fused_$tag.log
