"""Write-only HBM bandwidth on this GPU (cudaMemset and a torch fill), next to the sampler stage's store rate."""
import torch, json
x = torch.empty(503316480 // 8, dtype=torch.float64, device="cuda")
res = {}
for name, fn in (("memset_zero", lambda: x.zero_()), ("fill_kernel", lambda: x.fill_(1.5))):
    for _ in range(5): fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(20): fn()
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 20
    res[name] = {"ms": ms, "GB_per_s": x.numel() * 8 / ms / 1e6}
print(json.dumps(res))
