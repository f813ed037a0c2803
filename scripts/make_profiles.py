"""usage: python scripts/make_profiles.py [r2]
Turn the files scripts/round_end.sh left in gpurun_out/ into the tracked
artefacts under profiles/ (condensed launch list, launch shares, traffic.json,
ncu --set full summary, bench lines)."""
import collections, csv, io, json, shutil, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else 'r2'
rows = list(csv.reader(open('gpurun_out/%s_launches_raw.csv' % R)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
out = [['id', 'kernel', 'duration_us']]; agg = collections.defaultdict(list)
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    name = r[ki]; v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    short = name.split('(')[0]
    if 'xc::' in name or short.startswith('k_'):
        out.append([r[0], short, '%.2f' % v]); agg[short].append(v)
csv.writer(open('profiles/%s_final_launches.csv' % R, 'w')).writerows(out)
tot = sum(sum(v) for v in agg.values())
lines = ['kernel,launches,total_us,avg_us,share_of_xc_time']
for k, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
    lines.append('%s,%d,%.1f,%.2f,%.3f' % (k, len(v), sum(v), sum(v) / len(v), sum(v) / tot))
open('profiles/%s_final_launch_shares.csv' % R, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines[:8]))
raw = subprocess.run(['ncu', '-i', 'gpurun_out/prof_%s_final.ncu-rep' % R, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw))); h = rr[0]; u = rr[1]; idx = {x: i for i, x in enumerate(h)}
def val(r, k):
    v = float(r[idx[k]].replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'us': 1, 'ns': 1e-3, 'ms': 1e3}.get(u[idx[k]], 1)
tr = {}
for r in rr[2:]:
    n = r[idx['Kernel Name']]
    key = 'lwa' if ('k_lwa_cols' in n or 'k_lwa_fx<' in n or 'k_lwa_fast' in n) else 'bin_accumulate' if ('k_bin_rows' in n or 'k_hist' in n) else 'minmax_levels' if 'minmax' in n else 'epilogue' if 'epilogue' in n else None
    if key and key not in tr:
        grid = r[idx['launch__grid_size']]
        tr[key] = {'dram_bytes_read': val(r, 'dram__bytes_read.sum'), 'dram_bytes_write': val(r, 'dram__bytes_write.sum'),
                   'ncu_duration_us': val(r, 'gpu__time_duration.sum'), 'grid': grid}
flat = {k: v['dram_bytes_read'] + v['dram_bytes_write'] for k, v in tr.items()}
flat['_detail'] = tr
flat['_note'] = ('dram__bytes_read.sum + dram__bytes_write.sum per launch (one pass = 32 slices of 721x1440), '
                 'one ncu --set full capture (cold cache), scripts/round_end.sh, round ' + R)
json.dump(flat, open('profiles/traffic.json', 'w'), indent=1)
print({k: v for k, v in flat.items() if not k.startswith('_')})
for a in ('final_ncu_full_summary.txt', 'final_source_lines.txt', 'sanitizer.txt', 'bench_n1.json', 'bench_ref_n1.json', 'bench_c5.json', 'gpu_tests.txt'):
    try:
        shutil.copy('gpurun_out/%s_%s' % (R, a), 'profiles/%s_%s' % (R, a))
    except FileNotFoundError:
        print('missing', a)
d = json.loads(open('profiles/%s_bench_n1.json' % R).read().strip().splitlines()[-1]); r = d['roofline']
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
print('roofline', {k: r[k] for k in ('kernel', 'achieved', 'frac', 'traffic')}); print(r['stage_ms_per_step']); print(r['pipeline'])
print(d['clocks']); print(d['cpu_baseline']['value'])
