import torch, time
torch.cuda.init()
x = torch.empty(10*1024*1024//4, dtype=torch.float32).pin_memory()
y = torch.empty(87*1024*1024//40, dtype=torch.float32).pin_memory()
dx = torch.empty_like(x, device='cuda'); dy = torch.empty_like(y, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, n=20):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
def h2d():
    with torch.cuda.stream(s1): dx.copy_(x, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): y.copy_(dy, non_blocking=True)
def both(): h2d(); d2h()
a=t(h2d); b=t(d2h); c=t(both)
print("H2D %.1f MB: %.3f ms (%.1f GB/s)  D2H %.1f MB: %.3f ms (%.1f GB/s)  both: %.3f ms" % (x.numel()*4/1e6, a, x.numel()*4/a/1e6, y.numel()*4/1e6, b, y.numel()*4/b/1e6, c))

# write-combined pinned memory (cudaHostAllocWriteCombined = 4) vs torch's default pinned memory
import ctypes as C
rt = C.CDLL("libcudart.so.12")
n = x.numel() * 4
for flags, name in ((0, "default"), (4, "write-combined"), (1, "portable")):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags)) == 0
    C.memset(p, 1, n)
    st = torch.cuda.current_stream().cuda_stream
    def f():
        rt.cudaMemcpyAsync(C.c_void_p(dx.data_ptr()), p, C.c_size_t(n), C.c_int(1), C.c_void_p(st))
    ms = t(f)
    print("cudaHostAlloc(%s): H2D %.3f ms (%.1f GB/s)" % (name, ms, n / ms / 1e6))
    rt.cudaFreeHost(p)
