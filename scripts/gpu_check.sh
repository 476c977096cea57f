#!/bin/bash
# GPU contact of a change: dev check of the fused kernels (+ A/B builds), the gpu test-suite, the engine
# timeline, a bench run.
#   gpurun --timeout 1800 -- 'bash scripts/gpu_check.sh tag'
set -u
tag=${1:-chk}
out=gpurun_out
mkdir -p $out
timeout 600 python scripts/dev_fused.py > $out/dev_fused_$tag.log 2>&1; echo "dev_fused rc=$?"
# A/B: other builds of the same sources (development only, see auromat_b200/_lib.py AMT_LIB)
for v in auromat_b200/csrc/_variants/*.so; do
  [ -f "$v" ] || continue
  AMT_LIB=$PWD/$v timeout 600 python scripts/dev_fused.py > $out/dev_fused_${tag}_$(basename $v .so).log 2>&1; echo "dev_fused $v rc=$?"
done
grep -h "timings" $out/dev_fused_${tag}*.log | cut -c1-330
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 > $out/gputests_$tag.log 2>&1; echo "pytest rc=$?"
tail -5 $out/gputests_$tag.log
AMT_SEQ_TRACE=1 timeout 300 python scripts/seq_trace.py 40 > $out/seq_trace_$tag.txt 2>&1; echo "trace rc=$?"
grep -E "^==|^median" $out/seq_trace_$tag.txt
timeout 900 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
tail -3 $out/bench_$tag.err
python - <<PY
import json
d=json.loads(open('$out/bench_$tag.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print({k:(v.get('ms_per_step'), v.get('value')) for k,v in d['variants'].items()})
print('roofline', {k:d['roofline'][k] for k in ('kernel_ms','frac','achieved','peak')}, 'hbm', d['roofline']['hbm']['frac'])
print('parity', d.get('parity',{}).get('pass'))
PY
