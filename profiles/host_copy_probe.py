"""Probe of the host-side costs that bound the end-to-end path (run on the GPU box)."""
import time, numpy as np, torch, os
def t(f, n=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
print("cpus", os.cpu_count(), "THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
MB = 612
a = None
def fresh():
    global a; a = np.empty(MB << 20, dtype=np.uint8); a[::4096] = 1
print(f"first-touch {MB} MB: {t(fresh)*1e3:.1f} ms")
d = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
h_page = torch.from_numpy(a)
print(f"D2H pageable (touched): {t(lambda: h_page.copy_(d))*1e3:.1f} ms")
def fresh_copy():
    b = torch.empty(MB << 20, dtype=torch.uint8); b.copy_(d)
print(f"D2H pageable (fresh alloc): {t(fresh_copy)*1e3:.1f} ms")
h_pin = torch.empty(MB << 20, dtype=torch.uint8, pin_memory=True)
print(f"D2H pinned: {t(lambda: h_pin.copy_(d, non_blocking=True))*1e3:.1f} ms")
print(f"H2D pageable: {t(lambda: d.copy_(h_page))*1e3:.1f} ms;  H2D pinned: {t(lambda: d.copy_(h_pin, non_blocking=True))*1e3:.1f} ms")
src = h_pin.numpy(); dst = np.empty_like(src); dst[::4096] = 1
print(f"host memcpy touched {MB} MB: {t(lambda: np.copyto(dst, src))*1e3:.1f} ms")
def alloc_free():
    x = torch.cuda.caching_allocator_alloc(2_300_000_000); torch.cuda.caching_allocator_delete(x)
torch.cuda.empty_cache()
import ctypes
rt = ctypes.CDLL("libcudart.so.12")
p = ctypes.c_void_p()
def raw_malloc():
    rt.cudaMalloc(ctypes.byref(p), ctypes.c_size_t(2_300_000_000)); rt.cudaFree(p)
print(f"cudaMalloc+cudaFree 2.3 GB: {t(raw_malloc, 3)*1e3:.1f} ms")
