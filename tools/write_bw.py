"""Dev tool: pure-write HBM bandwidth (torch fill_ = a vectorised store kernel) for the streaming roofline."""
import torch
dev = torch.device("cuda:0")
for mb in (1024, 4096, 16384):
    buf = torch.empty(mb << 20, dtype=torch.uint8, device=dev).view(torch.float64)
    buf.fill_(1.0); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); buf.fill_(2.0); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    print(f"fill {mb} MiB: {(mb << 20) / min(ts) / 1e9:.1f} GB/s")
    del buf
