cat > /tmp/wf_once.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import radar_sounder_crw_b200 as crw
x = torch.randn(32, 10, 47, 128, device="cuda", requires_grad=True)
for _ in range(3):
    loss, _, _ = crw.ops.walk_loss(x, 0.07, False, crw.ops.PREC_BF16X3)
    loss.backward()
torch.cuda.synchronize()
print("ok", float(loss))
PY
python /tmp/wf_once.py > gpurun_out/wf_once.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:walk_fused -c 2 -o gpurun_out/r02d_ncu_walk_roles -f python /tmp/wf_once.py > gpurun_out/ncu_wf.log 2>&1
python bench.py --only walk --steps 2 --warmup 3 > gpurun_out/walk_only.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02d_launches_walk.csv python bench.py --only walk --steps 2 --warmup 3 > gpurun_out/ncu_wl.log 2>&1
ls -la gpurun_out/r02d*
