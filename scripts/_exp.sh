timeout 1200 python -m pytest tests -q -m gpu --timeout 900 -x 2>&1 | tail -5
timeout 300 python scripts/dev_fused.py 2>&1 | grep "iss timings" 
for m in 0 1 4 63; do echo "== skip mask $m"; AMT_LIB=$PWD/auromat_b200/csrc/_variants/devskip.so AMT_SEQ_SKIP=$m AMT_SEQ_TRACE=1 timeout 300 python scripts/seq_trace.py 80 2>&1 | grep -E "^median" | head -1; done
