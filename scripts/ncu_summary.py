"""Summarise ncu output into small text files for profiles/ (run in the build container).

    python scripts/ncu_summary.py report gpurun_out/prof.ncu-rep  > profiles/rNN_kernel.txt
    python scripts/ncu_summary.py launches gpurun_out/launches.csv > profiles/rNN_launches.txt
    python scripts/ncu_summary.py json gpurun_out/prof.ncu-rep "<command that was profiled>" > profiles/r02_ncu_kernels.json

The `json` form is what bench.py loads for `roofline.ncu` / `roofline.traffic` (no literals in bench.py):
per kernel the duration, DRAM bytes, pipe utilisation and executed-instruction counters, stamped with
the git commit of the tree the summary was made from.
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
    'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
]


def ncu(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def report(path):
    raw = list(csv.reader(io.StringIO(ncu(['-i', path, '--page', 'raw', '--csv']))))
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        name = row[hdr.index('Kernel Name')]
        print('== kernel:', name)
        for h, u, v in zip(hdr, units, row):
            if h in KEYS:
                print('  %-82s %18s %s' % (h, v, u))
    text = ncu(['-i', path, '--page', 'source', '--csv'])
    # the source page is one CSV table per kernel, each introduced by a "Kernel Name" row
    tables, cur = [], None
    for r in csv.reader(io.StringIO(text)):
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}
            tables.append(cur)
        elif cur is not None:
            cur['rows'].append(r)
    for t in tables:
        rows = t['rows']
        if not rows:
            continue
        h = rows[0]
        try:
            si, ei = h.index('Source'), h.index('Instructions Executed')
        except ValueError:
            continue
        agg, tot = collections.Counter(), 0
        for r in rows[1:]:
            if len(r) <= ei:
                continue
            try:
                n = int(float(r[ei]))
            except ValueError:
                continue
            toks = r[si].split()
            if not toks:
                continue
            op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
            agg[op] += n
            tot += n
        print('== executed warp instructions by SASS opcode: %s: total %d' % (t['name'][:60], tot))
        for k, v in agg.most_common(22):
            print('  %-10s %14d %5.1f%%' % (k, v, 100.0 * v / max(tot, 1)))


def to_json(paths, command=''):
    """`paths`: one report or several separated by commas (the first occurrence of a kernel wins)."""
    import json
    kernels = {}
    for path in paths.split(','):
        _json_kernels(path, kernels)
    import os
    git = os.environ.get('AMT_GIT') or subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], capture_output=True,
                                                      text=True).stdout.strip()
    print(json.dumps({'git': git, 'command': command, 'report': paths, 'kernels': kernels}, indent=1))


def _json_kernels(path, kernels):
    raw = list(csv.reader(io.StringIO(ncu(['-i', path, '--page', 'raw', '--csv']))))
    hdr = raw[0]
    short = {
        'gpu__time_duration.sum': 'duration_ns',
        'dram__bytes_read.sum': 'dram_bytes_read', 'dram__bytes_write.sum': 'dram_bytes_write',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active': 'pipe_fp64_pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active': 'pipe_alu_pct',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active': 'pipe_fma_pct',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active': 'pipe_lsu_pct',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active': 'pipe_xu_pct',
        'smsp__inst_executed.sum': 'warp_instructions',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum': 'thread_dfma',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum': 'thread_dmul',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum': 'thread_dadd',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_throughput_pct',
        'launch__registers_per_thread': 'registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
        'lts__t_sectors_op_red.sum': 'l2_red_sectors', 'lts__t_sectors_op_atom.sum': 'l2_atom_sectors',
    }
    units = raw[1]

    def scale(v, u):
        v = float(v.replace(',', ''))
        return v * {'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'us': 1e3, 'ms': 1e6, 'msecond': 1e6, 'usecond': 1e3,
                    'second': 1e9}.get(u, 1.0)
    for row in raw[2:]:
        name = row[hdr.index('Kernel Name')].replace('void ', '')
        if name in kernels:
            continue
        k = {}
        for h, u, v in zip(hdr, units, row):
            if h in short:
                try:
                    k[short[h]] = scale(v, u)
                except ValueError:
                    pass
        if 'dram_bytes_read' in k and 'dram_bytes_write' in k:
            k['dram_bytes'] = k['dram_bytes_read'] + k['dram_bytes_write']
        kernels[name] = k


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        agg.setdefault(r[ki], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print('%-72s %5s %12s %12s %7s' % ('kernel', 'n', 'avg_us', 'total_us', 'share'))
    for k, v in agg.items():
        print('%-72s %5d %12.1f %12.1f %6.1f%%' % (k[:72], len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3, 100 * sum(v) / tot))


if __name__ == '__main__':
    {'report': report, 'launches': launches, 'json': to_json}[sys.argv[1]](*sys.argv[2:])
