"""PCIe probe 3: does a write-combined upload buffer (cudaHostAllocWriteCombined) change the both-directions-at-once rate?"""
import ctypes as C, time, torch
torch.cuda.init()
rt = C.CDLL("libcudart.so.12")
nx, ny = 10 * 1024 * 1024, 8700 * 1024
dx = torch.empty(nx, dtype=torch.uint8, device="cuda"); dy = torch.empty(ny, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for flags, name in ((0, "default pinned"), (4, "write-combined X"), (1, "portable"), (2, "mapped")):
    px, py = C.c_void_p(), C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(px), C.c_size_t(nx), C.c_uint(flags)) == 0
    assert rt.cudaHostAlloc(C.byref(py), C.c_size_t(ny), C.c_uint(0)) == 0
    C.memset(px, 1, nx); C.memset(py, 1, ny)
    def h2d(): rt.cudaMemcpyAsync(C.c_void_p(dx.data_ptr()), px, C.c_size_t(nx), C.c_int(1), C.c_void_p(s1.cuda_stream))
    def d2h(): rt.cudaMemcpyAsync(py, C.c_void_p(dy.data_ptr()), C.c_size_t(ny), C.c_int(2), C.c_void_p(s2.cuda_stream))
    def both(): h2d(); d2h()
    def t(fn, n=40):
        fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
    a, b, c = t(h2d), t(d2h), t(both)
    print("%-18s H2D %.0f us (%.1f GB/s)  D2H %.0f us (%.1f GB/s)  both at once %.0f us" % (name, a, nx / a / 1e3, b, ny / b / 1e3, c))
    rt.cudaFreeHost(px); rt.cudaFreeHost(py)
