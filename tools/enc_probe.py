import torch, time, sys
sys.path.insert(0,'/root/repo')
import radar_sounder_crw_b200 as crw
torch.manual_seed(0)
x = torch.randn(15040,1,32,32, device='cuda')
def run(enc, x, steps=5, label=''):
    opt = torch.optim.Adam(enc.parameters(), lr=1e-3)
    for i in range(3):
        y = enc(x); l = y.float().square().mean(); opt.zero_grad(); l.backward(); opt.step()
    torch.cuda.synchronize(); t0=time.time()
    for i in range(steps):
        y = enc(x); l = y.float().square().mean(); opt.zero_grad(); l.backward(); opt.step()
    torch.cuda.synchronize(); print(label, (time.time()-t0)/steps*1e3, 'ms')
enc = crw.Resnet(False).cuda().train()
run(enc, x, label='fp32 default')
torch.backends.cudnn.benchmark=True
run(enc, x, label='fp32 cudnn.benchmark')
torch.backends.cuda.matmul.allow_tf32=True
enc2 = crw.Resnet(False).cuda().train().to(memory_format=torch.channels_last)
run(enc2, x.contiguous(memory_format=torch.channels_last), label='fp32 channels_last+benchmark+tf32')
with torch.autocast('cuda', dtype=torch.bfloat16):
    run(enc2, x.contiguous(memory_format=torch.channels_last), label='bf16 autocast channels_last')
