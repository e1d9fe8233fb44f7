"""peer access and copy bandwidth between GPU 0 and the others (run on a multi-GPU box)"""
import torch, time
n = torch.cuda.device_count()
print("gpus", n)
print("peer access from 0:", [torch.cuda.can_device_access_peer(0, j) for j in range(1, n)])
x = torch.empty(64 << 20, dtype=torch.uint8, device="cuda:0")
for j in range(1, n):
    y = torch.empty_like(x, device="cuda:%d" % j)
    for _ in range(2): y.copy_(x)
    torch.cuda.synchronize(0); torch.cuda.synchronize(j)
    t0 = time.perf_counter()
    for _ in range(10): y.copy_(x)
    torch.cuda.synchronize(0); torch.cuda.synchronize(j)
    print("0 -> %d: %.1f GB/s" % (j, 10 * x.numel() / (time.perf_counter() - t0) / 1e9))
