#!/usr/bin/env python
"""Per-kernel device time of one walk fwd+bwd on the tile engine (torch.profiler): B T N from argv (default 32 20 369)."""
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import radar_sounder_crw_b200 as crw  # noqa: E402

B, T, N = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (32, 20, 369)))
emb = torch.randn(B, T, N, 128, device="cuda", requires_grad=True)


def step():
    loss, _, _ = crw.ops.walk_loss(emb, 0.07, False, crw.ops.PREC_BF16X3)
    loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
tot, cnt = defaultdict(float), defaultdict(int)
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
for e in evs:
    tot[e.name] += e.device_time
    cnt[e.name] += 1
span = max(e.time_range.end for e in evs) - min(e.time_range.start for e in evs)
print(f"B={B} T={T} N={N}: {len(evs)} launches, sum {sum(tot.values()) / 1e3:.3f} ms, span {span / 1e3:.3f} ms")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v / 1e3:9.3f} ms  x{cnt[k]:<4d} {k[:110]}")
