"""pinned vs write-combined pinned host memory: host -> device bandwidth through cudart directly"""
import ctypes as C, time
rt = C.CDLL("libcudart.so")
n = 59 * 1024 * 1024
dev = C.c_void_p(); assert rt.cudaMalloc(C.byref(dev), C.c_size_t(2 * n)) == 0
st = [C.c_void_p(), C.c_void_p()]
for s in st:
    assert rt.cudaStreamCreate(C.byref(s)) == 0
for flags, name in ((0, "pinned (default)"), (4, "pinned write-combined")):
    hp = [C.c_void_p(), C.c_void_p()]
    for h in hp:
        assert rt.cudaHostAlloc(C.byref(h), C.c_size_t(n), C.c_uint(flags)) == 0
        C.memset(h, 1, n)
    for k in (1, 2):
        for rep in range(2):
            rt.cudaDeviceSynchronize(); t0 = time.perf_counter()
            for it in range(20):
                for j in range(2):
                    rt.cudaMemcpyAsync(C.c_void_p(dev.value + j * n), hp[j], C.c_size_t(n), C.c_int(1), st[j % k])
            rt.cudaDeviceSynchronize(); dt = time.perf_counter() - t0
        print(f"{name}, {k} stream(s): {40 * n / dt / 1e9:.2f} GB/s")
    for h in hp:
        rt.cudaFreeHost(h)
