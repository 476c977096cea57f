#!/bin/bash
# compute-sanitizer over the GPU tests that exercise this round's kernels (run through gpurun):
#   memcheck: fused kernel, limb solvers, sanitise tile, outline collect / eval, engine sequences, binning;
#   racecheck / synccheck: the kernels with shared memory and barriers (sanitise tile, fused SIP, limb SIP, outline).
#   gpurun --timeout 2400 -- 'bash scripts/gpu_sanitizer.sh r02'
set -u
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
K1='limb_solver or fused_kernel or sanitize_matches or plane_free or ring_planes or deterministic or config3 or sip_device or cell_indices or intersects_earth'
K2='sip_limb_solver or sanitize_matches or fused_kernel_equals or plane_free_fused'
compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest tests/test_gpu_parity.py -q -x -k "$K1" --timeout 3000 > $out/sanitizer_memcheck_$tag.log 2>&1
compute-sanitizer --tool racecheck --error-exitcode 0 python -m pytest tests/test_gpu_parity.py -q -x -k "$K2" --timeout 3000 > $out/sanitizer_racecheck_$tag.log 2>&1
compute-sanitizer --tool synccheck --error-exitcode 0 python -m pytest tests/test_gpu_parity.py -q -x -k "$K2" --timeout 3000 > $out/sanitizer_synccheck_$tag.log 2>&1
for t in memcheck racecheck synccheck; do echo "== $t"; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard" $out/sanitizer_${t}_$tag.log | tail -4; done
