#!/bin/bash
# Is this box's HBM as fast as the pool's usual one?  A plain device copy, then bench.py on the named libraries.
python - <<'PY'
import torch, time
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda"); y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(10):
    a.record(); y.copy_(x); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
print("copy 1 GiB: %.1f GB/s (read + write)" % (2 * (1 << 30) / best / 1e6))
z = torch.zeros(32, device="cuda")
for _ in range(100): z.add_(1.0)
torch.cuda.synchronize()
a.record()
for _ in range(2000): z.add_(1.0)
b.record(); torch.cuda.synchronize()
print("tiny kernel, back to back: %.2f us per launch" % (a.elapsed_time(b) / 2000 * 1e3))
w = torch.empty(680 << 20, dtype=torch.uint8, device="cuda"); v = torch.empty_like(w)
for _ in range(3): v.copy_(w)
torch.cuda.synchronize(); a.record()
for _ in range(300): v.copy_(w)
b.record(); torch.cuda.synchronize()
print("copy 680 MiB x 300 back to back: %.1f us per copy, %.1f GB/s" % (a.elapsed_time(b) / 300 * 1e3, 2 * (680 << 20) / (a.elapsed_time(b) / 300) / 1e6))
PY
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,temperature.memory --format=csv,noheader
for lib in "$@"; do
  if [ "$lib" = "default" ]; then unset VFK_LIB; else export VFK_LIB=$PWD/vfclik_b200/libvfk_$lib.so; fi
  timeout 300 python bench.py --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', '%.4f ms frac %.3f e2e %.3f' % (d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step']), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
