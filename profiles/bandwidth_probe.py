import torch, time
x = torch.empty(1<<30, dtype=torch.float64, device="cuda")  # 8 GiB
y = torch.empty(1<<29, dtype=torch.float64, device="cuda")
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
ms = t(lambda: x.zero_()); print("write-only (memset 8 GiB): %.0f GB/s" % (x.numel()*8/ms/1e6))
ms = t(lambda: x.fill_(1.5)); print("write-only (fill kernel 8 GiB): %.0f GB/s" % (x.numel()*8/ms/1e6))
ms = t(lambda: y.copy_(x[:1<<29])); print("copy 4 GiB -> 4 GiB: %.0f GB/s" % (2*y.numel()*8/ms/1e6))
ms = t(lambda: x.sum()); print("read-only (sum 8 GiB): %.0f GB/s" % (x.numel()*8/ms/1e6))
z = torch.empty(1<<29, dtype=torch.float64, device="cuda")
ms = t(lambda: torch.add(y, 1.0, out=z)); print("1 read : 1 write (4+4 GiB): %.0f GB/s" % (2*y.numel()*8/ms/1e6))
# 1 read : 5 writes like the assembly kernels
outs = [torch.empty(1<<27, dtype=torch.float64, device="cuda") for _ in range(5)]
src = torch.empty(1<<27, dtype=torch.float64, device="cuda")
def rw():
    for o in outs: o.copy_(src)
ms = t(rw); print("5 x (1 GiB read (L2-missing) -> 1 GiB write): %.0f GB/s" % (10*src.numel()*8/ms/1e6))
