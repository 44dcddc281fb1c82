import ctypes, torch, time
rt = ctypes.CDLL("libcudart.so.12")
n = 635_040_000
def alloc(flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(flags))
    assert rc == 0, rc
    buf = (ctypes.c_char * n).from_address(p.value)
    t = torch.frombuffer(buf, dtype=torch.float32)
    return t, p
torch.cuda.init()
d = torch.empty(n // 4, device="cuda")
for name, flags in (("default", 0), ("portable", 1), ("write-combined", 4)):
    t, p = alloc(flags)
    print(name, "is_pinned", t.is_pinned())
    t.copy_(torch.ones(1).expand(n // 4)[: n // 4]) if False else None
    for _ in range(2): d.copy_(t, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): d.copy_(t, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    print(name, "H2D GB/s", 5 * n / a.elapsed_time(b) / 1e6)
    rt.cudaFreeHost(p)
tp = torch.empty(n // 4, pin_memory=True)
for _ in range(2): d.copy_(tp, non_blocking=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): d.copy_(tp, non_blocking=True)
b.record(); torch.cuda.synchronize()
print("torch pinned H2D GB/s", 5 * n / a.elapsed_time(b) / 1e6)
