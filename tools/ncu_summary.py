"""Condense `ncu -i X.ncu-rep --page raw --csv` output into the per-launch summary tables kept under profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv
    python tools/ncu_summary.py gpurun_out/prof_raw.csv [-k regex] [--mean] [-c "comment line"] > profiles/rNx_ncu_summary.csv

One row per profiled launch (or, with --mean, per kernel name), the columns below when the capture holds them
(`--set full`); the second row carries ncu's units.  Metric columns are matched by suffix: newer ncu versions prefix
them with their collection pass (`SM_A.TriageCompute.…`)."""
import argparse
import csv
import re
import sys

METRICS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
]


def short_name(name):
    """`void dccf::k_x<2>(dccf::P)` -> `k_x<2>`"""
    name = re.sub(r'^void\s+', '', name)
    name = re.sub(r'\(.*$', '', name)
    return name.split('::')[-1] if '<' not in name else re.sub(r'^.*::(?=[^:<]*<)', '', name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('raw_csv')
    ap.add_argument('-k', '--kernel', default=None, help='regex on the kernel name')
    ap.add_argument('--mean', action='store_true', help='one row per kernel name: mean over its launches (+ count)')
    ap.add_argument('-c', '--comment', default=None)
    a = ap.parse_args()
    rows = [r for r in csv.reader(open(a.raw_csv, newline='')) if r]
    start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr, units, body = rows[start], rows[start + 1], rows[start + 2:]
    name_col = hdr.index('Kernel Name')
    cols = []
    for m in METRICS:
        hit = [i for i, h in enumerate(hdr) if h == m or h.endswith('.' + m)]
        if hit:
            cols.append((m, hit[0]))
    pat = re.compile(a.kernel) if a.kernel else None
    out = csv.writer(sys.stdout, lineterminator='\n')
    if a.comment:
        print('# ' + a.comment)
    out.writerow(['Kernel Name'] + [m for m, _ in cols] + (['launches'] if a.mean else []))
    out.writerow([''] + [units[i] for _, i in cols] + ([''] if a.mean else []))

    def num(x):
        try:
            return float(x.replace(',', ''))
        except ValueError:
            return None

    picked = [(short_name(r[name_col]), [r[i] for _, i in cols]) for r in body
              if len(r) > name_col and (pat is None or pat.search(r[name_col]))]
    if not a.mean:
        for name, vals in picked:
            out.writerow([name] + vals)
        return
    order, groups = [], {}
    for name, vals in picked:
        if name not in groups:
            groups[name] = []
            order.append(name)
        groups[name].append(vals)
    for name in order:
        g = groups[name]
        means = []
        for j in range(len(cols)):
            xs = [num(v[j]) for v in g]
            means.append('%.6g' % (sum(xs) / len(xs)) if all(x is not None for x in xs) else g[0][j])
        out.writerow([name] + means + [len(g)])


if __name__ == '__main__':
    main()
