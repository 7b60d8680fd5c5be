"""Host-to-device bandwidth from pinned memory on this box: one 128 MB copy vs column-chunked strided copies (torch).
The library itself chunks with cudaMemcpy2DAsync (47 GB/s effective inside pb_loglik_host, against 51.5 GB/s for one flat copy)."""
import torch, time
n=128; s=1_000_000
h=torch.empty((n,s),dtype=torch.uint8).pin_memory()
d=torch.empty((n,s),dtype=torch.uint8,device='cuda')
st=torch.cuda.Stream()
def t(f,reps=10):
    with torch.cuda.stream(st):
        f(); st.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps): f()
        e1.record(st); e1.synchronize()
    return e0.elapsed_time(e1)/reps
print('1D full   ', n*s/ t(lambda: d.copy_(h,non_blocking=True))/1e6,'GB/s')
for ch in (4,8,16):
    w=s//ch
    def f():
        for k in range(ch): d[:,k*w:(k+1)*w].copy_(h[:,k*w:(k+1)*w],non_blocking=True)
    print('2D chunks',ch, n*s/t(f)/1e6,'GB/s')
