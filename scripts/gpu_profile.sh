#!/bin/bash
# ncu evidence of a round (run on the B200 box through gpurun; summaries are made in the build
# container with scripts/ncu_summary.py and committed under profiles/).
#   gpurun --timeout 1200 -- 'bash scripts/gpu_profile.sh r02'
set -u
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
# 1. every launch of a short bench run with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 12 --warmup 3 --repeats 1 --no-variants --no-cpu-baseline > $out/bench_under_ncu_$tag.log 2>&1
# 2. full counter set of one launch of every kernel of the step (after warm-up): the long kernel and the
#    normalisation, then the stage-A kernels (they run three frames ahead of their fused kernel)
ncu --set full --clock-control none --import-source on \
    -k regex:'k_georef_fused|k_normalise' -s 40 -c 4 -f -o $out/step_$tag \
    python bench.py --steps 6 --warmup 3 --repeats 1 --no-variants --no-cpu-baseline > $out/ncu_step_$tag.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'k_limb_bits|k_sanitize|k_outline_collect|k_outline_eval' -s 24 -c 8 -f -o $out/stagea_$tag \
    python bench.py --steps 6 --warmup 3 --repeats 1 --no-variants --no-cpu-baseline > $out/ncu_stagea_$tag.log 2>&1
# 3. the scatter in three regimes: atomics / L2 counters of k_bin and k_georef_fused
ncu --set full --clock-control none -k regex:'k_bin|k_georef_fused' -f -o $out/scatter_$tag \
    python scripts/bin_scatter.py > $out/ncu_scatter_$tag.log 2>&1
REPS=10 python scripts/bin_scatter.py > $out/scatter_times_$tag.txt 2>&1
# 4. summaries are made HERE (gpurun brings back at most 64 MiB and a report with imported source is ~30 MB):
#    text + JSON of every report, then only the step report itself travels
python scripts/ncu_summary.py launches $out/launches_$tag.csv > $out/${tag}_launches.txt
python scripts/ncu_summary.py report $out/step_$tag.ncu-rep > $out/${tag}_step_ncu.txt
python scripts/ncu_summary.py report $out/stagea_$tag.ncu-rep > $out/${tag}_stagea_ncu.txt
python scripts/ncu_summary.py report $out/scatter_$tag.ncu-rep > $out/${tag}_scatter_ncu.txt
python scripts/ncu_summary.py json $out/step_$tag.ncu-rep,$out/stagea_$tag.ncu-rep,$out/scatter_$tag.ncu-rep \
    "bash scripts/gpu_profile.sh $tag" > $out/${tag}_ncu_kernels.json
ncu -i $out/step_$tag.ncu-rep --page source --csv --kernel-name regex:k_georef_fused --launch-count 1 > $out/${tag}_fused_source.csv 2>/dev/null
rm -f $out/stagea_$tag.ncu-rep $out/scatter_$tag.ncu-rep
ls -la $out | tail -16
