"""Executed warp-instructions and stall samples per CUDA source line of the first kernel in an .ncu-rep (needs -lineinfo + --import-source on).
usage: python scripts/ncu_lines.py <rep> [units] [top]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
per, samples, text = collections.Counter(), collections.Counter(), {}
fname, H, seen_fn = None, None, 0
for x in rows:
    if not x:
        continue
    if x[0] == "File Path":
        fname = x[1].split("/")[-1]; continue
    if x[0] == "Function Name":
        seen_fn += 1; continue
    if x[0] == "Line No":
        H = x; ie = H.index("Instructions Executed"); isamp = H.index("# Samples"); continue
    if H is None or x[0] == "":
        continue
    try:
        key = (fname, int(x[0])); per[key] += int(x[ie]); samples[key] += int(x[isamp]); text[key] = x[1].strip()
    except (ValueError, IndexError):
        pass
tot, ts = sum(per.values()), sum(samples.values())
print("total warp-instructions %d = %.0f per unit; samples %d" % (tot, tot / units, ts))
for key, n in per.most_common(top):
    print("%6.1f/unit %5.1f%% inst %5.1f%% samp  %s:%d  %s" % (n / units, 100.0 * n / tot, 100.0 * samples[key] / max(ts, 1), key[0], key[1], text[key][:90]))
