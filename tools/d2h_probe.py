"""D2H rate for result-sized copies (6.4 MB) into pinned memory, CUDA-event timed."""
import torch
for mb in (0.5, 6.4, 64.0):
    n = int(mb * 1e6)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    for _ in range(3): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("D2H %.1f MB: %.3f ms = %.1f GB/s" % (mb, ms, n / ms / 1e6))
