"""Developer tool (GPU): dense int8 tensor-core peak through torch._int_mm (cuBLASLt), burst and sustained."""
import torch, time
n = 8192
a = torch.randint(-128, 127, (n, n), dtype=torch.int8, device="cuda")
b = torch.randint(-128, 127, (n, n), dtype=torch.int8, device="cuda").t().contiguous().t()   # column-major B
for _ in range(3):
    torch._int_mm(a, b)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch._int_mm(a, b); e1.record(); e1.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("int8 burst  %.1f TOP/s (%.3f ms)" % (2 * n ** 3 / best / 1e9, best))
reps = max(10, int(1.5e3 / best))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    torch._int_mm(a, b)
e1.record(); e1.synchronize()
print("int8 sustained %.1f TOP/s over %d reps" % (2 * n ** 3 * reps / e0.elapsed_time(e1) / 1e9, reps))
