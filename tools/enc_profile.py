"""One-off: where does the (PyTorch, out-of-scope) encoder spend the train step?  Prints the top CUDA kernels."""
import sys
import torch
sys.path.insert(0, '/root/repo')
import radar_sounder_crw_b200 as crw
from torch.profiler import profile, ProfilerActivity
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
seq = torch.randn(32, 10, 47, 32, 32, device='cuda')
enc = crw.Resnet(False).cuda().train().to(memory_format=torch.channels_last)
model = crw.CRW(enc, 0.07, False, need_A=False)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
def step():
    loss, _ = model(seq); opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
