"""Brief of an .ncu-rep: headline launch metrics, stall reasons per issue and the dynamic opcode mix of the first kernel.
usage: python scripts/ncu_brief.py gpurun_out/<name>.ncu-rep [units_per_launch]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, r = rows[0], rows[2]
def g(name):
    return r[H.index(name)] if name in H else "n/a"
print("kernel:", g("Kernel Name")[:110])
for m in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
          "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"):
    print("  %-70s %s" % (m, g(m)))
st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[i]))
      for i, h in enumerate(H) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
print("  stalls per issue:", ", ".join("%s %.2f" % x for x in sorted(st, key=lambda x: -x[1])[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
blocks = [i for i, x in enumerate(rows) if x and x[0] == "Kernel Name"]
end = blocks[1] if len(blocks) > 1 else len(rows)
H = rows[1]
ia, ie = H.index("Source"), H.index("Instructions Executed")
ops, tot = collections.Counter(), 0
for x in rows[2:end]:
    try:
        n = int(x[ie])
    except (ValueError, IndexError):
        continue
    t = x[ia].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0].rstrip(";")
    ops[op] += n
    tot += n
print("  warp-instructions executed: %d%s" % (tot, (" = %.0f per unit" % (tot / units)) if units else ""))
print("  opcode mix:", ", ".join("%s %.1f%%" % (o, 100.0 * n / tot) for o, n in ops.most_common(22)))
