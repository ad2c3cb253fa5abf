"""Summarise an .ncu-rep of render_kernel: key metrics, stall mix, per-region instruction counts per path.
usage: python tools/ncu_summary.py report.ncu-rep paths_in_launch"""
import collections, csv, io, re, subprocess, sys
rep, paths = sys.argv[1], float(sys.argv[2])
def page(p):
    out = subprocess.run(["ncu", "-i", rep, "--page", p, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))
rows = page("raw"); hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'dram__bytes_write.sum', 'dram__bytes_read.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__warps_eligible.avg.per_cycle_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for h, u, v in zip(hdr, units, vals):
    if h in keep or ('issue_stalled' in h and 'per_issue_active' in h and float(v) > 0.05):
        print(f"{h} [{u}] = {v}")
rows = page("source"); hdr = rows[1]; data = rows[2:]
ia, isrc, ismp, it, noi = (hdr.index(k) for k in ('Instructions Executed', 'Source', '# Samples', 'Avg. Threads Executed', 'stall_no_inst'))
tot = sum(float(r[ia]) for r in data); totS = sum(float(r[ismp]) for r in data)
print(f"SASS instrs {len(data)}, executed {tot:.3g} = {tot / paths:.1f} warp-instr per path")
chunk = 250
for c in range(0, len(data), chunk):
    seg = data[c:c + chunk]
    ex = sum(float(r[ia]) for r in seg); sm = sum(float(r[ismp]) for r in seg)
    if ex / tot < 0.005: continue
    thr = sum(float(r[ia]) * float(r[it]) for r in seg) / max(ex, 1)
    ops = collections.Counter(re.sub(r'\..*', '', r[isrc].split()[0] if not r[isrc].startswith('@') else r[isrc].split()[1]) for r in seg)
    ni = sum(float(r[noi]) for r in seg)
    print(f"{c:5d}: exec {100 * ex / tot:5.1f}% ({ex / paths:6.1f}/path) samples {100 * sm / totS:5.1f}% noinst {100 * ni / max(sm, 1):3.0f}% thr {thr:4.1f}  "
          + ", ".join(f"{k}:{v}" for k, v in ops.most_common(6)))
