"""Substitute the @PLACEHOLDERS@ of DESIGN.md with the numbers of a bench.py JSON line (and a --config c5 line).
usage: python scripts/fill_docs.py profiles/r2_bench_n1.json [profiles/r2_bench_c5.json]"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
c5 = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]) if len(sys.argv) > 2 else None
r = d["roofline"]; st = r["stage_ms_per_step"]; fr = r["stage_frac"]
sub = {"@VALUE@": "%.1f" % (d["value"] / 1e3), "@LWA_MS@": "%.3f" % st["lwa"], "@LWA_FRAC@": "%.1f" % (100 * fr["lwa"]),
       "@BIN_MS@": "%.3f" % st["bin_accumulate"], "@BIN_FRAC@": "%.1f" % (100 * fr["bin_accumulate"]),
       "@MM_MS@": "%.3f" % st["minmax_levels"], "@MM_FRAC@": "%.0f" % (100 * fr["minmax_levels"]),
       "@PIPE_FRAC@": "%.1f" % (100 * r["pipeline"]["frac"]), "@E2E@": "%.1f" % (d["e2e"]["value"] / 1e3)}
if c5:
    rc = c5["roofline"]
    sub["@C5@"] = "%.3f ms/slice = %.0f GB/s = **%.1f %%** (`k_bin_rows`, in-flight Cartesian stencil); %.0f slices/s for the whole Keff part" % (
        rc["stage_ms_per_step"]["bin_accumulate"] / c5["config"]["slices_per_step_per_gpu"], rc["achieved"], 100 * rc["frac"], c5["value"])
s = open("DESIGN.md").read()
for k, v in sub.items():
    s = s.replace(k, v)
open("DESIGN.md", "w").write(s)
print(sub)
