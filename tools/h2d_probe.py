import torch, time
n = 1 << 28  # 1 GiB of float32
h = torch.empty(n, dtype=torch.float32, pin_memory=True); d = torch.empty(n, dtype=torch.float32, device='cuda')
def t(fn, k=5):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/k
one = t(lambda: d.copy_(h, non_blocking=True))
s = [torch.cuda.Stream() for _ in range(4)]
def multi(m):
    per = n // m
    for i in range(m):
        with torch.cuda.stream(s[i]): d[i*per:(i+1)*per].copy_(h[i*per:(i+1)*per], non_blocking=True)
print('1 stream GB/s', 4*n/one/1e9)
for m in (2,4): print(m,'streams GB/s', 4*n/t(lambda: multi(m))/1e9)
