"""Pure-write HBM bandwidth probe (torch fill_ / cudaMemset) to put k_render's store stream in context."""
import torch
dev = torch.device("cuda")
for mb in (906, 2048):
    x = torch.empty(mb * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    for name, fn in (("fill_", lambda: x.fill_(1.0)), ("zero_", lambda: x.zero_())):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"{name} {mb} MiB: {best*1e3:.1f} us -> {x.numel()*4/best/1e6:.1f} GB/s")
    y = torch.empty_like(x)
    for _ in range(3): y.copy_(x)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"copy_ {mb} MiB: {best*1e3:.1f} us -> {2*x.numel()*4/best/1e6:.1f} GB/s (read+write)")
