"""Measure write-only (memset) and copy bandwidth on this GPU for the buffer sizes the front end writes."""
import torch
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1e-3
for mb in (281, 1024, 4096):
    n = mb*1024*1024//4
    bufs=[torch.empty(n, device='cuda') for _ in range(3)]
    i=[0]
    def z():
        bufs[i[0]%3].zero_(); i[0]+=1
    def f():
        bufs[i[0]%3].fill_(1.5); i[0]+=1
    src=torch.randn(n, device='cuda')
    def c():
        bufs[i[0]%3].copy_(src); i[0]+=1
    def rd():
        bufs[i[0]%3].sum(); i[0]+=1
    tz=t(z); tf=t(f); tc=t(c); tr=t(rd)
    print(f"{mb} MB: zero_ {n*4/tz/1e9:.0f} GB/s  fill_ {n*4/tf/1e9:.0f} GB/s  copy(r+w) {2*n*4/tc/1e9:.0f} GB/s  sum(read) {n*4/tr/1e9:.0f} GB/s")
