import os, sys
sys.path.insert(0, "/root/repo")
import torch
import radar_sounder_crw_b200 as crw
T, N, M = 1250, 49, 4
g = torch.Generator().manual_seed(1)
feats = torch.randn(1, T, N, 128, generator=g).pin_memory()
mask0 = torch.nn.functional.one_hot(torch.randint(0, M, (1, N), generator=g), M).permute(0, 2, 1).float().contiguous().cuda()
flush = torch.empty(64 << 20, device="cuda")
def step(prec):
    l, _, _, _ = crw.ops.labelprop_host(feats, mask0, 20, 12.0, 0.07, 10, 0, True, False, prec)
    return l.cpu()
for name, prec in [("exact", crw.ops.PREC_TC_EXACT), ("bf16x3", crw.ops.PREC_BF16X3)]:
    for _ in range(3): step(prec)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(prec); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sum(ts) / len(ts)
    print(f"{name}: {ms:.3f} ms  {20000 / ms / 1e3:.2f} M columns/s  (env SEGS={os.environ.get('CRW_LP_HOST_SEGS')}, SEG_SMS={os.environ.get('CRW_LP_HOST_SEG_SMS')})")
