"""pinned host -> device bandwidth with 1, 2 and 4 concurrent copy streams (is one DMA queue enough to saturate the link?)"""
import torch, time
n = 59 * 1024 * 1024
host = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(4)]
dev = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(4)]
for k in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(k)]
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for it in range(20):
            for i, s in enumerate(streams):
                with torch.cuda.stream(s):
                    for j in range(i, 4, k):
                        dev[j].copy_(host[j], non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{k} stream(s): {20 * 4 * n / dt / 1e9:.2f} GB/s")
