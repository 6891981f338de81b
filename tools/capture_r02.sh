set -x
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
LP_PREC=tcx python tools/lp_once.py > gpurun_out/lp_once_tcx.log 2>&1 && ncu --nvtx --nvtx-include "lp_call/" --metrics $M --clock-control none --csv --log-file gpurun_out/r02b_ncu_lp_call_tcx.csv python tools/lp_once.py > gpurun_out/ncu_tcx.log 2>&1
LP_PREC=tcx LP_CFG=5 python tools/lp_once.py > gpurun_out/lp_once_c5.log 2>&1 && LP_PREC=tcx LP_CFG=5 ncu --nvtx --nvtx-include "lp_call/" --metrics $M --clock-control none --csv --log-file gpurun_out/r02b_ncu_lp_call_cfg5_tcx.csv python tools/lp_once.py > gpurun_out/ncu_c5.log 2>&1
CRW_LP_NO_FORK=1 ncu --set full --clock-control none --import-source on -k regex:"lp_filter|lp_refine|lp_prep_x" -c 3 -o gpurun_out/r02b_ncu_lp_x -f python tools/lp_once.py > gpurun_out/ncu_x.log 2>&1
cat > /tmp/wf_once.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
os.environ["CRW_WALK_FUSED"] = "1"
import torch
import radar_sounder_crw_b200 as crw
x = torch.randn(32, 10, 47, 128, device="cuda", requires_grad=True)
for _ in range(3):
    loss, _, _ = crw.ops.walk_loss(x, 0.07, False, crw.ops.PREC_BF16X3)
    loss.backward()
torch.cuda.synchronize()
print("ok", float(loss))
PY
python /tmp/wf_once.py > gpurun_out/wf_once.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:walk_fused -c 2 -o gpurun_out/r02b_ncu_walk_fused -f python /tmp/wf_once.py > gpurun_out/ncu_wf.log 2>&1
python bench.py --only labelprop --steps 2 --warmup 3 > gpurun_out/lp_only.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_lp.csv python bench.py --only labelprop --steps 2 --warmup 3 > gpurun_out/ncu_lpl.log 2>&1
python bench.py --only walk --steps 2 --warmup 3 > gpurun_out/walk_only.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02b_launches_walk.csv python bench.py --only walk --steps 2 --warmup 3 > gpurun_out/ncu_wl.log 2>&1
ls -la gpurun_out | tail -12
