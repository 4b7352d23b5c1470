for c in 1 2 4 8 16 32; do VFK_SESSION_CHUNKS=$c python bench.py --steps 50 --no-cpu-baseline --no-extras | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunks $c', 'e2e %.3e'%d['e2e']['value'], 'ms %.3f'%d['e2e']['ms_per_step'])"; done
python - <<'PY'
import torch, time
n=29360128
a=torch.empty(n,dtype=torch.uint8).pin_memory(); b=torch.empty(n,dtype=torch.uint8).pin_memory()
d=torch.empty(n,dtype=torch.uint8,device='cuda'); e=torch.empty(n,dtype=torch.uint8,device='cuda')
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def run(both):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(20):
        with torch.cuda.stream(s1): d.copy_(a,non_blocking=True)
        if both:
            with torch.cuda.stream(s2): b.copy_(e,non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter()-t)/20
run(True)
print("H2D only: %.3f ms (%.1f GB/s)"%(run(False)*1e3, n/run(False)/1e9))
t=run(True); print("H2D+D2H concurrent: %.3f ms (%.1f GB/s each way)"%(t*1e3, n/t/1e9))
PY
