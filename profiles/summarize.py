#!/usr/bin/env python3
"""Turn ncu outputs in gpurun_out/ into the small text summaries committed under profiles/.
  python profiles/summarize.py launches gpurun_out/launches_r1a.csv          > profiles/r1a_launches.txt
  python profiles/summarize.py raw      gpurun_out/prof_scatter_r1a.ncu-rep  > profiles/r1a_scatter_raw.txt
  python profiles/summarize.py source   gpurun_out/prof_refine_r1c.ncu-rep k_refine > profiles/r1c_refine_source.txt
"""
import csv, subprocess, sys

RAW = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
       'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
       'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
       'launch__occupancy_limit_registers', 'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
       'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
       'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def launches(path):
    rows = [r for r in csv.reader(open(path, errors='replace')) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = {}
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(',', ''))
        except ValueError:
            continue
        name = r[ik].split('(')[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    unit = rows[1][hdr.index('Metric Unit')]
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.0f} {unit} in kernels (cold-cache, serialised: compare SHARES)")
    print(f"{'kernel':28s} {'launches':>8s} {'time':>12s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:28s} {n:8d} {t:12.0f} {t / tot:7.3f}")


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    print(f"# {path} (ncu --set full --clock-control none), one block per captured launch")
    for r in rows[2:]:
        print(r[hdr.index('Kernel Name')].split('(')[0])
        for m in RAW:
            if m in hdr:
                print(f"  {m:75s} {r[hdr.index(m)]:>16s} {rows[1][hdr.index(m)]}")


def source(path, kernel):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', kernel,
                          '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No'][0]
    hdr = rows[hi]
    ii, iw = hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
    res = []
    for r in rows[hi + 1:]:
        if r and r[0].strip().isdigit():
            try:
                res.append((int(r[0]), r[1], float(r[ii]), float(r[iw])))
            except ValueError:
                pass
    ti, tw = sum(x[2] for x in res), sum(x[3] for x in res)
    print(f"# {path} {kernel}: {ti:.0f} warp instructions, {tw:.0f} stall samples; top source lines by stall share")
    for ln, src, v, w in sorted(res, key=lambda x: -x[3])[:25]:
        print(f"{ln:5d} inst {100 * v / ti:5.1f}% stall {100 * w / tw:5.1f}%  {src.strip()[:110]}")


if __name__ == '__main__':
    {'launches': launches, 'raw': raw, 'source': source}[sys.argv[1]](*sys.argv[2:])
