import os, sys
sys.path.insert(0, "/root/repo")
import torch
from torch.profiler import profile, ProfilerActivity
import radar_sounder_crw_b200 as crw
T, N, M = 1250, 49, 4
g = torch.Generator().manual_seed(1)
feats = torch.randn(1, T, N, 128, generator=g).pin_memory()
mask0 = torch.nn.functional.one_hot(torch.randint(0, M, (1, N), generator=g), M).permute(0, 2, 1).float().contiguous().cuda()
def step():
    l, _, _, _ = crw.ops.labelprop_host(feats, mask0, 20, 12.0, 0.07, 10, 0, True, False, crw.ops.PREC_TC_EXACT)
    return l.cpu()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
evs = sorted((e for e in prof.events() if e.device_type.name == "CUDA"), key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
for e in evs:
    print(f"{e.name[:50]:50s} start {e.time_range.start - t0:8.1f} us  dur {e.time_range.end - e.time_range.start:8.1f} us")
