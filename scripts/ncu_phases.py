"""Executed warp-instructions per line of the KERNEL BODY (call sites), each SASS address counted once and attributed to the
kernel-body line it was inlined into.  usage: python scripts/ncu_phases.py <rep> <units> [first_body_line]"""
import collections, csv, subprocess, sys
rep, units = sys.argv[1], float(sys.argv[2])
body0 = int(sys.argv[3]) if len(sys.argv) > 3 else 555
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
addr_lines, addr_n, addr_op, text = collections.defaultdict(list), {}, {}, {}
fname, H, cur, nfn = None, None, None, 0
for x in rows:
    if not x: continue
    if x[0] == "File Path": fname = x[1].split("/")[-1]; continue
    if x[0] == "Function Name":
        nfn += 1; continue
    if x[0] == "Line No":
        H = x; ia = H.index("Address"); ie = H.index("Instructions Executed"); continue
    if H is None: continue
    if x[0] != "":
        cur = (fname, int(x[0])); text[cur] = x[1].strip(); continue
    a = x[ia]
    try: n = int(x[ie])
    except ValueError: continue
    addr_lines[a].append(cur); addr_n[a] = n; addr_op[a] = x[ia + 1].split()
per, ops = collections.Counter(), collections.defaultdict(collections.Counter)
for a, ls in addr_lines.items():
    body = [l for l in ls if l[0] == "vfk_kernels.cuh" and l[1] >= body0]
    key = max(body, key=lambda l: l[1]) if body else ls[-1]
    per[key] += addr_n[a]
    t = addr_op[a]; op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0].rstrip(";")
    ops[key][op] += addr_n[a]
tot = sum(per.values())
print("total %.0f per unit" % (tot / units))
for key in sorted(per, key=lambda k: (k[0] != "vfk_kernels.cuh", k[1])):
    if per[key] / units < 0.5: continue
    print("%7.1f  %s:%d  %-70s | %s" % (per[key] / units, key[0], key[1], text[key][:70], " ".join("%s:%.0f" % (o, n / units) for o, n in ops[key].most_common(6))))
