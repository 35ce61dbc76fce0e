"""Summarises an .ncu-rep: per kernel key metrics + top stalled source lines.  usage: ncu_summary.py file.ncu-rep [kernel-regex] [n_lines]
   ncu_summary.py --traffic out.json stage=file.ncu-rep:kernel-substring ...   writes dram read + write bytes per launch of the named kernels (bench.py roofline.traffic)"""
import csv, io, json, subprocess, sys
if len(sys.argv) > 2 and sys.argv[1] == "--traffic":
    out = {}
    for spec in sys.argv[3:]:
        stage, rest = spec.split("=", 1)
        rep, sub = rest.rsplit(":", 1)
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]
        ki, ri, wi = hdr.index('Kernel Name'), hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
        unit = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        vals = [float(r[ri]) * unit[rows[1][ri]] + float(r[wi]) * unit[rows[1][wi]] for r in rows[2:] if sub in r[ki]]
        if vals:
            out[stage] = sum(vals) / len(vals)
    open(sys.argv[2], "w").write(json.dumps(out, indent=1) + "\n")
    print(out)
    sys.exit(0)
rep = sys.argv[1]; rx = sys.argv[2] if len(sys.argv) > 2 else None; nl = int(sys.argv[3]) if len(sys.argv) > 3 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
ki = hdr.index('Kernel Name')
for r in rows[2:]:
    print("=====", r[ki][:60])
    for w in want:
        if w in hdr:
            i = hdr.index(w); print(f"   {w:72s} {r[i]:>16s} {units[i]}")
if rx:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + rx], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    his = [i for i, r in enumerate(rows) if len(r) > 5 and 'Address' in r]
    for n, hi in enumerate(his[:1]):
        h = rows[hi]; end = his[n + 1] if n + 1 < len(his) else len(rows); data = [r for r in rows[hi + 1:end] if len(r) > 10]
        si, sc, ie = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
        cols = [c for c in ('stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_barrier', 'stall_math', 'stall_lg', 'stall_mio', 'stall_not_selected') if c in h]
        tot = sum(int(r[si] or 0) for r in data)
        print(f"--- {rx}: {len(data)} SASS instrs, {tot} samples; columns: samples, executed, " + ", ".join(c[6:] for c in cols))
        for r in sorted(data, key=lambda r: -int(r[si] or 0))[:nl]:
            print(r[si].rjust(6), r[ie].rjust(9), " ".join(r[h.index(c)].rjust(4) for c in cols), " ", r[sc][:100])
