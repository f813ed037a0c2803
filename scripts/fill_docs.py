"""Write DESIGN.md and profiles/README.md from their .template files, substituting the @PLACEHOLDERS@ with the numbers
of the round's bench.py JSON lines under profiles/ (so that the prose never carries a number no run produced).
usage: python scripts/fill_docs.py [round tag, default r2]"""
import json, os, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line(name):
    p = os.path.join(ROOT, "profiles", "%s_%s.json" % (tag, name))
    if not os.path.exists(p):
        return None
    return json.loads(open(p).read().strip().splitlines()[-1])


d, c5, ref = line("bench_n1"), line("bench_c5"), line("bench_ref_n1")
r = d["roofline"]; st = r["stage_ms_per_step"]; fr = r["stage_frac"]
sub = {"@VALUE@": "%.1f" % (d["value"] / 1e3), "@MS@": "%.3f" % d["ms_per_step"],
       "@LWA_MS@": "%.3f" % st["lwa"], "@LWA_FRAC@": "%.1f" % (100 * fr["lwa"]), "@LWA_GB@": "%.0f" % r["achieved"],
       "@BIN_MS@": "%.3f" % st["bin_accumulate"], "@BIN_FRAC@": "%.1f" % (100 * fr["bin_accumulate"]),
       "@MM_MS@": "%.3f" % st["minmax_levels"], "@MM_FRAC@": "%.0f" % (100 * fr["minmax_levels"]),
       "@EPI_MS@": "%.3f" % st["epilogue"],
       "@PIPE_FRAC@": "%.1f" % (100 * r["pipeline"]["frac"]), "@E2E@": "%.1f" % (d["e2e"]["value"] / 1e3),
       "@C2D@": "%.0f" % d["e2e_contour2d"]["value"] if (d.get("e2e_contour2d") or {}).get("value") else "n/a",
       "@CPU@": "%.2f" % d["cpu_baseline"]["value"], "@CORES@": "%d" % d["cpu_baseline"]["cores"]}
if ref:
    sub["@REF@"] = "%.2f" % ref["value"]
if c5:
    rc = c5["roofline"]; n5 = c5["config"]["slices_per_step_per_gpu"]
    sub["@C5@"] = "%.3f ms/slice = %.0f GB/s = **%.1f %%** (`k_bin_rows`, in-flight Cartesian stencil); %.0f slices/s for the whole Keff part" % (
        rc["stage_ms_per_step"]["bin_accumulate"] / n5, rc["achieved"], 100 * rc["frac"], c5["value"])
    sub["@C5_MS@"] = "%.3f" % (rc["stage_ms_per_step"]["bin_accumulate"] / n5)
    sub["@C5_EPI_MS@"] = "%.3f" % (rc["stage_ms_per_step"]["epilogue"] / n5)
    sub["@C5_FRAC@"] = "%.1f" % (100 * rc["frac"]); sub["@C5_VAL@"] = "%.0f" % c5["value"]
# the N >= 2 lines are compared with the N = 1 line of the SAME library build (the multi-GPU runs were made before the
# last kernel change of the round; profiles/<tag>_bench_n1_scaling_base.json is the N = 1 line of that build)
base = line("bench_n1_scaling_base") or d
sub["@V1@"] = "%.0f" % base["value"]; sub["@MS1@"] = "%.3f" % base["ms_per_step"]; sub["@E1@"] = "%.0f" % base["e2e"]["value"]
d_final, d = d, base
for n in (2, 4, 8):
    dn = line("bench_n%d" % n)
    if dn:
        sub["@V%d@" % n] = "%.0f" % dn["value"]; sub["@MS%d@" % n] = "%.3f" % dn["ms_per_step"]
        sub["@EF%d@" % n] = "%.3f" % (dn["value"] / (n * d["value"])); sub["@E%d@" % n] = "%.0f" % dn["e2e"]["value"]
for src, dst in (("DESIGN.md.template", "DESIGN.md"), ("profiles/README.md.template", "profiles/README.md")):
    s = open(os.path.join(ROOT, src)).read()
    for k, v in sub.items():
        s = s.replace(k, v)
    import re
    left = sorted(set(re.findall(r"@[A-Z0-9_]+@", s)))
    if left:
        print("unfilled in %s: %s" % (dst, left))
    open(os.path.join(ROOT, dst), "w").write(s)
print(sub)
