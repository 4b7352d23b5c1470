"""Turn gpurun_out/<tag>_prof.ncu-rep + <tag>_launches.csv into the small text files committed under profiles/.

usage: python scripts/summarise_profile.py <tag> [workload]
"""
import csv, json, os, subprocess, sys
tag = sys.argv[1]
workload = sys.argv[2] if len(sys.argv) > 2 else "config3"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
os.makedirs(pr, exist_ok=True)
for name in ("launches.csv", "bench.json"):
    src = os.path.join(go, "%s_%s" % (tag, name))
    if os.path.exists(src):
        open(os.path.join(pr, "%s_%s" % (tag, name)), "w").write(open(src).read())
rep = os.path.join(go, tag + "_prof.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ("Kernel Name", "gpu__time_duration", "dram__bytes", "dram_throughput", "registers_per_thread", "occupancy",
        "inst_executed.sum", "issue_active", "warps_active", "stalled", "pipe_fma", "pipe_xu", "pipe_alu", "pipe_lsu",
        "lts__t_bytes.sum", "sm__throughput", "thread_inst_executed_per_inst", "shared_mem", "l1tex__t_bytes.sum",
        "Grid Size", "Block Size", "sass_thread_inst_executed_op", "cycles_elapsed.avg.per_second", "pipe_fp64")
keep = [i for i, h in enumerate(hdr) if any(k in h for k in keys)]
with open(os.path.join(pr, tag + "_ncu_full_summary.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + ["launch%d" % k for k in range(len(data))])
    for i in keep:
        w.writerow([hdr[i], units[i]] + [r[i] for r in data])
def col(name):
    i = hdr.index(name)
    return [float(r[i].replace(",", "")) for r in data], units[i]
rd, u1 = col("dram__bytes_read.sum")
wr, u2 = col("dram__bytes_write.sum")
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
traffic = sum(a * scale[u1] + b * scale[u2] for a, b in zip(rd, wr)) / len(rd)
tj = os.path.join(pr, "traffic.json")
d = json.load(open(tj)) if os.path.exists(tj) else {}
d[workload] = traffic
d[workload + "_source"] = "%s_ncu_full_summary.csv: dram__bytes_read.sum + dram__bytes_write.sum, mean of %d launches" % (tag, len(rd))
sys.path.insert(0, root)
import bench  # noqa: E402  (kernel_source_stamp: the sources this capture was taken on)
d[workload + "_stamp"] = bench.kernel_source_stamp()
json.dump(d, open(tj, "w"), indent=1)
dur, du = col("gpu__time_duration.sum")
dur = [x * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(du, 1.0) for x in dur]      # microseconds
print("traffic per launch %.1f MB, duration %s us" % (traffic / 1e6, dur))


# executed FLOP/s from the SASS op counters ("achieved FLOP/s ... from ncu against both peaks")
def rate(op):
    name = "smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op
    return col(name)[0] if name in hdr else [0.0] * len(data)
try:
    clk, cu = col("smsp__cycles_elapsed.avg.per_second")
    clk = [c * {"Ghz": 1e9, "Mhz": 1e6, "hz": 1.0}.get(cu, 1e9) for c in clk]
    f32 = [(a + m + 2 * f) * c / 1e12 for a, m, f, c in zip(rate("fadd"), rate("fmul"), rate("ffma"), clk)]
    f64 = [(a + m + 2 * f) * c / 1e12 for a, m, f, c in zip(rate("dadd"), rate("dmul"), rate("dfma"), clk)]
    pk = json.load(open(os.path.join(pr, "r01_pipe_peaks.json")))
    d[workload + "_executed_flops"] = {
        "fp32_tflops": sum(f32) / len(f32), "fp64_tflops": sum(f64) / len(f64),
        "fp32_frac_of_measured_ffma_peak": sum(f32) / len(f32) / pk["fp32_ffma_tflops"],
        "fp64_frac_of_measured_dfma_peak": sum(f64) / len(f64) / pk["fp64_dfma_tflops"],
        "hbm_tbs": traffic / (sum(dur) / len(dur) * 1e-6) / 1e12,
        "source": "%s_ncu_full_summary.csv: smsp__sass_thread_inst_executed_op_{f,d}{add,mul,fma}_pred_on (FMA = 2)" % tag}
    json.dump(d, open(tj, "w"), indent=1)
    print("executed: FP32 %.2f TFLOP/s, FP64 %.2f TFLOP/s" % (sum(f32) / len(f32), sum(f64) / len(f64)))
except Exception as exc:
    print("no executed-FLOP summary:", exc)
