#!/bin/bash
# Multi-GPU record of a round (run through `gpurun --gpus N`): copy-only H2D ceiling and bench.py
# (frame-sharded sequence + the configs[4] mosaic all-reduce variant) at every rank count up to N.
#   gpurun --gpus 8 --timeout 1500 -- 'bash scripts/gpu_scale.sh r02 8'
set -u
tag=${1:-r02}
nmax=${2:-2}
out=gpurun_out
mkdir -p $out
port=29510
for n in 1 2 4 8; do
  [ $n -le $nmax ] || continue
  port=$((port+1))
  if [ $n -eq 1 ]; then
    timeout 300 python scripts/h2d_ceiling.py >> $out/h2d_ceiling_$tag.txt 2>$out/h2d_err_$n.txt
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
        scripts/h2d_ceiling.py >> $out/h2d_ceiling_$tag.txt 2>$out/h2d_err_$n.txt
  fi
done
cat $out/h2d_ceiling_$tag.txt
for n in 2 4 8; do
  [ $n -le $nmax ] || continue
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n > $out/bench_${tag}_n$n.json 2> $out/bench_${tag}_n$n.err
  echo "bench n=$n rc=$?"
  tail -c 600 $out/bench_${tag}_n$n.err
  python - <<PY
import json
try:
    d=json.loads(open('$out/bench_${tag}_n$n.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('n_gpus','value','ms_per_step')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
    print('mosaic', d['variants'].get('mosaic'))
except Exception as e:
    print('no bench line', e)
PY
done
