"""Per-source-line instruction and stall-sample shares of one kernel from an ncu report taken with
--import-source on (profiles/r1_final_source_lines.txt).
usage: python scripts/ncu_source_lines.py <kernel regex> [top N] [report.ncu-rep]"""
import subprocess, sys
kern = sys.argv[1]
rep = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/prof_r1_final.ncu-rep"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
cur = None; data = []; hdr = None
for line in out.splitlines():
    f = line.strip()[1:-1].split('","')
    if len(f) == 2 and f[0] in ("File Path", "File Name"): cur = f[1].split("/")[-1]; continue
    if f and f[0] == "Line No": hdr = f; L = len(f); ii = hdr.index("Instructions Executed") - L; si = hdr.index("# Samples") - L; continue
    if hdr is None or len(f) < 10 or not f[0].isdigit(): continue
    try:
        data.append((cur, int(f[0]), f[1], int(f[ii] or 0), int(f[si] or 0)))
    except ValueError:
        pass
tot_i = sum(x[3] for x in data); tot_s = sum(x[4] for x in data)
print("total warp-instr", tot_i, "samples", tot_s)
data.sort(key=lambda x: -x[3])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for fn, ln, src, ins, smp in data[:n]:
    print("%5.1f%% ins %5.1f%% smp  %s:%d  %s" % (100.0 * ins / tot_i, 100.0 * smp / max(tot_s, 1), fn, ln, src.strip()[:105]))
