"""Summarise an .ncu-rep (raw + source pages) into text: key metrics per kernel,
stall-reason shares and the hottest SASS instructions.  Usage:
    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [topN]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 18
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.avg',
        'sm__inst_executed_pipe_fp64.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("=" * 100); print(r[idx['Kernel Name']][:110])
    for w in want:
        if w in idx: print("  %-68s %s %s" % (w, r[idx[w]], units[idx[w]]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
ker = None; hdr = None; data = {}
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == 'Kernel Name': ker = r[1][:70]; data.setdefault(ker, []); continue
    if r and r[0] == 'Address': hdr = r; continue
    if ker and hdr and len(r) > 10: data[ker].append(r)
for ker, d in data.items():
    si = hdr.index('# Samples'); ii = hdr.index('Instructions Executed')
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[si]) for r in d) or 1
    print("=" * 100); print(ker, " samples", tot, " warp-instr", sum(int(r[ii]) for r in d))
    agg = {s: sum(int(r[hdr.index(s)]) for r in d) for s in stalls}
    print("  stalls: " + ", ".join("%s %.1f%%" % (s[6:], 100 * v / tot) for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    for r in sorted(d, key=lambda r: -int(r[si]))[:topn]:
        top = max(stalls, key=lambda s: int(r[hdr.index(s)]))
        print("  %6d %5.1f%%  exec %9s  %-58s %s" % (int(r[si]), 100 * int(r[si]) / tot, r[ii], r[1].strip()[:58], top[6:]))
