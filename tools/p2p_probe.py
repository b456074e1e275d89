"""Dev tool: peer-copy bandwidth between cuda:0 and cuda:1 (the ceiling for the learner-buffer gather)."""
import torch, time
a = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
b = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:1")
print("can access peer:", torch.cuda.can_device_access_peer(0, 1))
for size in (13 << 20, 256 << 20):
    for _ in range(3):
        b[:size].copy_(a[:size])
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    t0 = time.perf_counter()
    for _ in range(10):
        b[:size].copy_(a[:size])
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    dt = (time.perf_counter() - t0) / 10
    print(f"copy {size >> 20} MiB: {dt * 1e6:.1f} us, {size / dt / 1e9:.1f} GB/s")
