"""Attribute executed instructions of a kernel to CUDA source lines (needs -lineinfo + --import-source on).
    python scripts/ncu_lines.py rep.ncu-rep kernel_substring [topN]"""
import collections, csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
cur_file = cur_fn = None
agg = collections.defaultdict(int); srcs = {}; seen_fn = None
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name':
        cur_fn = r[1]; continue
    if r[0] == 'Line No' or pat not in (cur_fn or ''): continue
    if r[0] != '' and len(r) > 7:
        try: n = int(float(r[7]))
        except ValueError: continue
        key = (cur_file, int(r[0])); agg[key] += n; srcs[key] = r[1].strip()[:110]
tot = sum(agg.values())
print('kernel ~ %s: %d executed warp instructions attributed (all captured launches)' % (pat, tot))
for key, n in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    print('  %6.2f%%  %-18s %s' % (100.0 * n / tot, '%s:%d' % key, srcs[key]))
