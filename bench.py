#!/usr/bin/env python
"""bench.py -- CRW hot-path benchmark (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus 8 --steps 20 --warmup 5
    python bench.py --impl reference --steps 3 --warmup 1       # CPU arm (oracle port) on the host cores

Prints ONE JSON line on rank 0.  Primary workload = BASELINE config 2: CRW train step, B=32 per GPU,
T=10 frames, N=47 nodes (400 rows, 32x32 patches, overlap 24), ResNet encoder (PyTorch), fused CUDA
walk fwd+bwd, Adam step.  The same line carries a "labelprop" object = BASELINE config 3: label
propagation over one 400 x 20k-column radargram per GPU (T=1250, N=49, 4 classes, ctx 20, k 10, r 12).

value      whole-job radargrams/s (columns/s for labelprop) with inputs resident in HBM
e2e        same, through the public API from pinned HOST buffers, H2D + D2H inside the timed region
roofline   the path's own CUDA kernels: algorithmic FLOPs (walk; tensor bound) or bytes (labelprop; hbm
           bound) over CUDA-event time, against MEASURED_PEAKS.json
cpu_baseline  oracle port on the box's host cores, bounded sample (rank 0, N=1 only)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

TRAIN = dict(B=32, T=10, H=400, patch=(32, 32), overlap=(24, 0), tau=0.07, lr=1e-3)       # config 2
TRAIN_CPU_SAMPLE_B = 2
LP = dict(R=1, rows=400, cols=20000, patch=(16, 16), overlap=(8, 0), M=4, ctx=20, k=10, radius=12, temp=0.07, C=128)
LP5 = dict(R_total=64, rows=400, cols=50000, patch=(16, 16), overlap=(8, 0), M=4, ctx=20, k=20, radius=24, temp=0.07, C=128)
COLS_PER_FRAME = 16


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_RESULT_FD = None


def protect_stdout():
    """Libraries print to stdout behind our back (NCCL: "NCCL version ..." at init).  Keep the original stdout for the ONE
    JSON line and point fd 1 at stderr for everybody else."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML) -- runs DURING the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.stop_flag, self.max_mhz = [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            log("nvml unavailable:", e)
            self.nv = None

    def _run(self):
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv:
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *exc):
        self.stop_flag = True
        if self.nv:
            self.t.join()
        return False

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def dist_env():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    return rank, world, local


def barrier_sync(world):
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    import torch.distributed as dist
    if world == 1:
        return x
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bind_to_gpu_numa(local, world):
    """Multi-GPU runs: pin this rank's host threads to the NUMA node of its GPU BEFORE any pinned buffer is allocated (first
    touch places the pages there), so that eight ranks do not stream their host features through one socket.  Returns what was
    found, for the bench line."""
    info = dict(numa_node=None, cpus=None, bound=False)
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/numa_node"
        node = int(open(path).read().strip())
        info["numa_node"] = node
        if node >= 0 and world > 1:
            cpus = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info.update(cpus=len(cpus), bound=True)
    except Exception as e:  # noqa: BLE001
        info["error"] = str(e)[:80]
    return info


def h2d_rate(world, mb=256):
    """Pinned host -> device copy rate of every rank with all ranks copying at once (GB/s): the floor of any end-to-end number
    that starts from host buffers."""
    import torch.distributed as dist
    src = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    dst = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    dst.copy_(src, non_blocking=True)
    barrier_sync(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbs = 4 * (mb << 20) / (e0.elapsed_time(e1) * 1e-3) / 1e9
    if world == 1:
        return [round(gbs, 1)]
    t = torch.tensor([gbs], device="cuda", dtype=torch.float64)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [round(float(o.item()), 1) for o in out]


class L2Flusher:
    """Writes a buffer larger than the 126 MB L2 between timed iterations (outside the CUDA-event brackets)."""

    def __init__(self, mb=256):
        self.buf = torch.empty(mb * 1024 * 1024 // 4, device="cuda", dtype=torch.float32)

    def __call__(self):
        self.buf.add_(1.0)


def timed_loop(step_fn, steps, warmup, world, flush=None, sampler=None):
    """W untimed warm-ups, then K steps, each bracketed by CUDA events on the launching stream; the whole loop
    is bracketed by barrier + synchronize.  Returns mean ms/step (max over ranks)."""
    for i in range(warmup):
        step_fn(i)
    barrier_sync(world)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ctx = sampler if sampler is not None else _Null()
    with ctx:
        torch.cuda.nvtx.range_push("timed")       # lets `ncu --nvtx --nvtx-include "timed/"` list the timed launches only
        for i in range(steps):
            if flush is not None:
                flush()
            ev[i][0].record()
            step_fn(warmup + i)
            ev[i][1].record()
        barrier_sync(world)
        torch.cuda.nvtx.range_pop()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    return max_over_ranks(ms, world)


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def synth_train_batch(B, T, seed):
    """randn radargrams [B, 400, T*32] cut exactly as the reference dataset does (src/dataset.py:34-39)."""
    g = torch.Generator().manual_seed(seed)
    (h, w), (oh, ow) = TRAIN["patch"], TRAIN["overlap"]
    n = (TRAIN["H"] - oh) // (h - oh)                                   # dataset.py:22
    rg = torch.randn(B, TRAIN["H"], T * w - ow * (T - 1), generator=g)
    item = rg[:, : n * h - oh * (n - 1)].unfold(1, h, h - oh).unfold(2, w, w - ow)   # [B,N,T,h,w]
    return item.permute(0, 2, 1, 3, 4).contiguous().float()


def walk_flops(B, T, N, C):
    f_aff = B * (T - 1) * 2 * N * N * C
    f_walk = B * max(3 * T - 10, 0) * 2 * N ** 3
    return 3 * (f_aff + f_walk)       # fwd + bwd


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def bench_train(crw, args, rank, world, local, pk, cfg4=False):
    """BASELINE config 2: full train step (PyTorch encoder + fused CUDA walk fwd/bwd + Adam).
    cfg4: BASELINE config 4 -- T=20 frames and the UNet-as-encoder adapter (UNet(1,128) + global average pool; NOT reference
    behaviour, see encoder.UNetEncoder), B=32 per GPU, all-reduce of the 17.2 MB of gradients."""
    import torch.distributed as dist  # noqa: F401
    B, T, tau = TRAIN["B"], (20 if cfg4 else TRAIN["T"]), TRAIN["tau"]
    steps = max(2, min(args.steps, 4)) if cfg4 else args.steps
    warm = 2 if cfg4 else args.warmup
    n_rot = 2 if cfg4 else 4   # rotate distinct input batches: 4 x 61.6 MB (2 x 123 MB) > 126 MB L2
    batches_host = [synth_train_batch(B, T, 1000 * rank + i).pin_memory() for i in range(n_rot)]
    batches_dev = [b.cuda() for b in batches_host]
    N = batches_dev[0].shape[2]
    torch.manual_seed(11)
    # the encoder is plain PyTorch (out of scope as a kernel); channels_last + TF32 are host-side settings
    torch.backends.cudnn.benchmark = not os.environ.get("CRW_BENCH_NO_AUTOTUNE")    # off only to keep ncu launch lists short
    if cfg4:
        encoder = crw.UNetEncoder(pos_embed=False, chunk=2048).cuda().train().to(memory_format=torch.channels_last)
    else:
        encoder = crw.Resnet(pos_embed=False).cuda().train().to(memory_format=torch.channels_last)
    # the walk runs on the fused tcgen05 kernels (precision=BF16X3: error-compensated bf16 pairs, fp32 accumulate -- BASELINE config 2's
    # "bf16 walk"; loss 1e-7 / gradients 1e-5 of the fp64 oracle)
    model = crw.CRW(encoder, tau, False, need_A=False, precision=crw.ops.PREC_BF16X3)
    fg = None
    if world > 1 and os.environ.get("CRW_BENCH_DDP"):
        # DistributedDataParallel, ~4 buckets of the 19.9 MB of gradients; no per-step broadcast of the batch-norm running statistics
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True,
                                                          bucket_cap_mb=int(os.environ.get("CRW_BENCH_BUCKET_MB", "5")),
                                                          broadcast_buffers=bool(int(os.environ.get("CRW_BENCH_DDP_BCAST", "0"))))
    # train-loop glue (scripts/train.py:56,70-72).  optim.FlatAdam: parameters, gradients and moments as views of flat
    # buffers, zero_grad = one memset, step = one SUM all-reduce of the gradient buffer + ONE launch of crw_b200::adam_step over all
    # parameters (the 1 / world rides in the kernel).  Default for N > 1: 34.71 ms per step on 2 B200s against 34.82 ms for
    # parallel.FlatGradients + torch.optim.Adam(fused=True).  At N = 1 torch's optimizer on fresh gradients stays the default
    # (34.53 against 34.66 ms: autograd ACCUMULATES into gradient views that already exist, one add per parameter, where
    # zero_grad(set_to_none) lets it hand the fresh gradient over).  CRW_BENCH_FLAT_ADAM=0 / 1 forces one or the other.
    use_flat = (world > 1) if os.environ.get("CRW_BENCH_FLAT_ADAM") is None else bool(int(os.environ["CRW_BENCH_FLAT_ADAM"]))
    use_flat = use_flat and not os.environ.get("CRW_BENCH_DDP")
    flat = None
    if use_flat:
        opt = flat = crw.optim.FlatAdam(model.parameters(), lr=TRAIN["lr"], world=world)
    else:
        if world > 1 and not os.environ.get("CRW_BENCH_DDP"):
            # the parameters' gradients are views of ONE flat buffer, averaged by one NCCL all-reduce after backward
            # (parallel.FlatGradients): no hooks, no buckets, no copies -- measured against DDP in DESIGN.md section 5
            fg = crw.parallel.FlatGradients(model.parameters(), world)
        opt = torch.optim.Adam(model.parameters(), lr=TRAIN["lr"], fused=True)
    last_loss = [None]

    def train_step(i, src=batches_dev):
        seq = src[i % n_rot]
        if not seq.is_cuda:
            seq = seq.cuda(non_blocking=True)
        loss, _ = model(seq)
        if flat is not None:
            flat.zero_grad()
            loss.backward()
        elif fg is not None:
            fg.zero()
            loss.backward()
            fg.reduce()
        else:
            opt.zero_grad(set_to_none=True)
            loss.backward()
        opt.step()
        last_loss[0] = loss

    sampler = ClockSampler(local)
    ms = timed_loop(train_step, steps, warm, world, sampler=sampler)

    # end to end from pinned HOST batches: a double-buffered input pipeline (what any data loader does) -- while step i
    # computes, the H2D copy of batch i+1 runs on a copy stream into the other device buffer; every timed step still
    # contains exactly one full batch copy and the D2H read of its loss.
    copy_stream = torch.cuda.Stream()
    dev_buf = [torch.empty_like(batches_dev[0]) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def enqueue_copy(i):
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])                 # the step that read this buffer has finished with it
            dev_buf[slot].copy_(batches_host[i % n_rot], non_blocking=True)
            ready[slot].record(copy_stream)

    for ev in consumed:
        ev.record()
    enqueue_copy(0)

    def train_step_e2e(i):
        slot = i % 2
        enqueue_copy(i + 1)                                        # next step's input, overlapped with this step
        torch.cuda.current_stream().wait_event(ready[slot])
        loss, _ = model(dev_buf[slot])
        if flat is not None:
            flat.zero_grad()
            loss.backward()
            consumed[slot].record()
        elif fg is not None:
            fg.zero()
            loss.backward()
            consumed[slot].record()
            fg.reduce()
        else:
            opt.zero_grad(set_to_none=True)
            loss.backward()
            consumed[slot].record()
        opt.step()
        last_loss[0] = loss.item()                                 # D2H read of the step's result

    # timed_loop numbers its steps warmup + i: keep the slot sequence continuous across warm-up and timed steps
    e2e_counter = [0]

    def train_step_e2e_seq(_i):
        train_step_e2e(e2e_counter[0])
        e2e_counter[0] += 1

    ms_e2e = timed_loop(train_step_e2e_seq, steps, 2 if cfg4 else 3, world)
    return dict(steps=steps, warmup=warm, frames=T, ms_per_step=ms, value=B * world / (ms * 1e-3), N=N, clocks=sampler.summary(), loss=float(last_loss[0]),
                optimizer=("optim.FlatAdam (crw_b200::adam_step, one launch over all parameters)" if flat is not None else "torch.optim.Adam(fused=True)"),
                e2e=dict(value=B * world / (ms_e2e * 1e-3), unit="radargrams/s",
                         h2d_bytes_per_step=int(batches_host[0].numel() * 4), d2h_bytes_per_step=4))


def bench_walk(crw, args, world, pk, N=47, T=None, B=None, precision=None, kernel_note=None, launches=8):
    """The hot path alone: fused walk fwd + bwd on resident embeddings (config-2 shape by default)."""
    B = TRAIN["B"] if B is None else B
    T = TRAIN["T"] if T is None else T
    tau = TRAIN["tau"]
    PREC = crw.ops.PREC_FP32 if precision is None else precision
    emb = torch.randn(B, T, N, 128, device="cuda", requires_grad=True)

    def walk_step(i):
        loss, _, _ = crw.ops.walk_loss(emb, tau, False, PREC)
        emb.grad = None
        loss.backward()

    ms_eager = timed_loop(walk_step, max(args.steps, 20), 5, world)
    # the same fwd+bwd captured once in a CUDA graph: removes the Python / custom-op dispatch gaps between the launches
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            loss, _, _ = crw.ops.walk_loss(emb, tau, False, PREC)
            torch.autograd.grad(loss, emb)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss, _, _ = crw.ops.walk_loss(emb, tau, False, PREC)
        gemb, = torch.autograd.grad(loss, emb)
    ms = timed_loop(lambda i: graph.replay(), max(args.steps, 20), 5, world)
    fl = walk_flops(B, T, N, 128)
    tf = fl / (ms * 1e-3) / 1e12
    kernel = kernel_note or ("walk fwd+bwd kernels (fp32, shared-memory-resident small-N path), 8 launches" if N <= 64 else
                             "walk fwd+bwd kernels (fp32 FMA tiles), 8 launches")
    return dict(ms=ms, ms_eager=ms_eager, launches=launches, shape=dict(B=B, T=T, N=N, C=128),
                roofline=dict(bound="tensor", achieved=tf, peak=pk["bf16"], unit="TFLOP/s", frac=tf / pk["bf16"],
                              traffic=None, kernel=kernel, algorithmic_flops=fl, peak_source=pk["src"] + " bf16 burst",
                              note=("N=47: launch/latency bound; tensor-pipe ceiling at this N is <= ~36% (SURVEY 7.3.1)"
                                    if N <= 64 else "scaled geometry (SURVEY appendix D), not a reference-size figure")))


def bench_walk_sweep(crw, args, world, pk):
    """Walk fwd+bwd on both engines at the reference size and at the scaled geometries of SURVEY appendix D."""
    out = []
    tcn = "walk fwd+bwd kernels, tcgen05 bf16x3 GEMMs (3 passes; algorithmic FLOPs counted once), 8 launches"
    for (B, T, N) in [(32, 10, 47), (32, 20, 47), (32, 20, 185), (32, 20, 369)]:
        for prec, note in [(crw.ops.PREC_FP32, None), (crw.ops.PREC_BF16X3, tcn)]:
            r = bench_walk(crw, args, world, pk, N=N, T=T, B=B, precision=prec, kernel_note=note)
            r["precision"] = "fp32" if prec == crw.ops.PREC_FP32 else "bf16x3"
            out.append(r)
            log(f"walk B={B} T={T} N={N} {r['precision']}: {r['ms']:.3f} ms  {r['roofline']['achieved']:.2f} TFLOP/s")
    return out


def bench_labelprop(crw, args, rank, world, pk, lp_config=None):
    """BASELINE config 3: one 400 x 20k-column radargram per GPU, features -> labels (weak scaling).
    lp_config 5: BASELINE config 5, 64 radargrams of 400 x 50k columns sharded over the ranks (strong scaling)."""
    global LP
    cfg5 = (args.lp_config if lp_config is None else lp_config) == 5
    lp_saved = LP
    steps = max(2, min(args.steps, 5)) if cfg5 else args.steps
    if cfg5:
        from radar_sounder_crw_b200.parallel import shard_range
        b, e = shard_range(LP5["R_total"], rank, world)
        LP = dict(LP5, R=e - b)
    Tl = LP["cols"] // COLS_PER_FRAME
    Nl = (LP["rows"] - LP["overlap"][0]) // (LP["patch"][0] - LP["overlap"][0])
    R, C, M = LP["R"], LP["C"], LP["M"]
    g = torch.Generator().manual_seed(77 + rank)
    feats_host = torch.randn(R, Tl, Nl, C, generator=g).pin_memory()
    label0 = torch.randint(0, M, (R, Nl), generator=g)
    mask0 = torch.stack([crw.utils.one_hot_mask(label0[r], M) for r in range(R)]).cuda()
    feats_dev = feats_host.cuda()
    res = [None]

    prec = [crw.ops.PREC_BF16X3]

    def lp_step(i, src=feats_dev):
        f = src if src.is_cuda else src.cuda(non_blocking=True)
        labels, _, _, _ = crw.ops.labelprop(f, mask0, LP["ctx"], float(LP["radius"]), LP["temp"], LP["k"],
                                            crw.ops.LP_REF_EXACT, prec[0], True, False)
        res[0] = labels

    flush = L2Flusher()
    fp32_path = None
    labels32 = None
    if args.lp_precision in ("both", "fp32") and not cfg5:
        prec[0] = crw.ops.PREC_FP32
        ms32 = timed_loop(lp_step, steps, args.warmup, world, flush=flush)
        labels32 = res[0].clone()
        fp32_path = dict(ms_per_step=ms32, value=R * Tl * COLS_PER_FRAME * world / (ms32 * 1e-3), unit="columns/s",
                         note="fp32 FMA path, pinned order, bit-exact against oracle/crw_oracle.c")
    total_cols = (LP5["R_total"] if cfg5 else R * world) * Tl * COLS_PER_FRAME

    def lp_step_e2e(i):
        # from pinned HOST features to labels on the host.  The public host-buffer calls stream the features in pieces, the copy of piece
        # c+1 overlapping the search of piece c: chunks of query tiles for the bf16x3 kernel, whole radargrams for the exact path (one
        # radargram alone is copied first: a segment's search lasts one filter item however short it is); fp32 copies first
        if prec[0] == crw.ops.PREC_BF16X3 or (prec[0] == crw.ops.PREC_TC_EXACT and R > 1):
            labels, _, _, _ = crw.ops.labelprop_host(feats_host, mask0, LP["ctx"], float(LP["radius"]), LP["temp"], LP["k"],
                                                     crw.ops.LP_REF_EXACT, True, False, prec[0])
            res[0] = labels
        else:
            lp_step(i, feats_host)
        res[0] = res[0].cpu()                   # D2H of the labels

    # PRIMARY: the exact tensor path (one tcgen05 fp16 pass filters with a proven margin, survivors re-scored in fp32 in the
    # pinned order): the only tensor-core path whose results are the fp32 path's on every input, near-collinear encoder
    # features included (tests/test_gpu_hard_cases.py).  The error-compensated bf16 kernel of round 1 is timed beside it.
    bf16x3_path = None
    if args.lp_precision != "fp32":
        prec[0] = crw.ops.PREC_BF16X3
        msb = timed_loop(lp_step, steps, args.warmup, world, flush=flush)
        labels_b = res[0].clone()
        msb_e2e = timed_loop(lp_step_e2e, steps, 3, world, flush=flush)
        bf16x3_path = dict(ms_per_step=msb, value=total_cols / (msb * 1e-3), unit="columns/s",
                           e2e=dict(value=total_cols / (msb_e2e * 1e-3), unit="columns/s", note="host-streamed entry (chunked H2D overlapped with the top-k)"),
                           kernel="lp_prep_bf16 + lp_topk_tc_kernel (tcgen05 kind::f16 on bf16 hi/lo pairs, 3 passes) + gather kernels",
                           note="precision=BF16X3 (round 1): approximate -- 100 % label agreement on these features, 99.86 .. 100 % on "
                                "near-collinear encoder features (tests/test_gpu_hard_cases.py)")
        prec[0] = crw.ops.PREC_TC_EXACT
    else:
        prec[0] = crw.ops.PREC_FP32
    ms = timed_loop(lp_step, steps, args.warmup, world, flush=flush)
    identical = None
    if labels32 is not None and args.lp_precision != "fp32":
        identical = bool((labels32 == res[0]).all().item())
        fp32_path["labels_identical_to_primary"] = identical
    if bf16x3_path is not None:
        bf16x3_path["label_agreement_with_primary"] = float((labels_b == res[0]).float().mean().item())
    ms_e2e = timed_loop(lp_step_e2e, steps, 3, world, flush=flush)
    pipeline = None
    if not cfg5 and args.lp_precision != "fp32":
        pipeline = bench_lp_pipeline(crw, args, world, flush, Tl, Nl, M)
    LP = lp_saved
    cols = R * Tl * COLS_PER_FRAME
    if cfg5:
        cols = LP5["R_total"] * Tl * COLS_PER_FRAME / world     # value below multiplies by world
    lp_bytes = R * Tl * (Nl * C * 4 + Nl * 4)   # read features once + write labels
    gbs = lp_bytes / (ms * 1e-3) / 1e9
    cfgd = LP5 if cfg5 else lp_saved
    dense = R * Tl * (cfgd["ctx"] + 1) * Nl * Nl * C * 2
    traffic = lp_call_traffic(cfg5, args.lp_precision)
    return dict(
        metric="labelprop_columns_per_sec", unit="columns/s", value=cols * world / (ms * 1e-3), ms_per_step=ms,
        e2e=dict(value=cols * world / (ms_e2e * 1e-3), unit="columns/s",
                 h2d_bytes_per_step=int(feats_host.numel() * 4), d2h_bytes_per_step=int(R * Tl * Nl * 4)),
        roofline=dict(bound="hbm", achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"],
                      # dram__bytes_read.sum + dram__bytes_write.sum summed over ALL kernels of one call (ncu capture named in
                      # traffic_source; see lp_call_traffic)
                      traffic=(traffic["bytes_per_radargram"] * R if traffic["bytes_per_radargram"] else None), traffic_source=traffic["source"],
                      kernel="lp_prep_x + lp_filter_kernel (tcgen05 kind::f16, one pass) + lp_refine_kernel (fp32 chain dots) + gather kernels"
                      if args.lp_precision != "fp32" else "l2_normalize + lp_topk_f32_kernel + gather kernels", algorithmic_bytes=lp_bytes,
                      peak_source=pk["src"],
                      tensor_bound_note=f"dense {dense / 1e9:.1f} GFLOP: AI ~514 FLOP/B > ridge, see DESIGN.md"),
        # the physically binding roofline (AI above the ridge): algorithmic dense FLOPs over the whole step; the error-compensated
        # bf16 path executes three MMA passes, counted once here
        roofline_tensor=dict(bound="tensor", achieved=dense / (ms * 1e-3) / 1e12, peak=pk["bf16"], unit="TFLOP/s",
                             frac=dense / (ms * 1e-3) / 1e12 / pk["bf16"], algorithmic_flops=dense,
                             note="dense (ctx+1) N^2 C 2 per query frame; 41 % of it lies inside the radius band; executed once (fp16 filter)"),
        gpu_launches=7 * steps, steps=steps, bf16x3_path=bf16x3_path, bit_exact=True, pipeline_e2e=pipeline,
        dtype="fp16 tensor-core filter + fp32 refine (results = fp32 pinned order)" if args.lp_precision != "fp32" else "f32",
        frames=Tl, fp32_path=fp32_path,
        scaling="strong" if cfg5 else "weak",
        config=dict(workload=(f"BASELINE config 5: 64 radargrams of 400x50000 columns sharded over {world} GPU(s) "
                              f"({R} on this rank) -> T=3125 x N=49 x C=128, M=4, ctx=20, k=20, radius=24, temp=0.07" if cfg5 else
                              "BASELINE config 3: 400x20000-column radargram per GPU -> T=1250 frames x N=49 nodes x C=128, "
                              "M=4, ctx=20, k=10, radius=12, temp=0.07, mode=ref_exact"),
                    l2="flushed between iterations (256 MB write)"))


def bench_lp_pipeline(crw, args, world, flush, Tl, Nl, M):
    """The reference's real call pattern for one radargram (scripts/test/test_all.py:91-96): host radargram + host segmentation ->
    RGDataset (HBM) -> frames -> encoder (PyTorch Resnet, eval) -> propagate -> nearest upsample -> label map back on the host.
    The encoder is out of scope as a kernel and dominates; reported so that the hot path's end-to-end weight is on record."""
    H, Wc = LP["rows"], LP["cols"]
    g = torch.Generator().manual_seed(5)
    rg_host = torch.randn(H, Wc, generator=g).pin_memory()
    seg_host = torch.randint(0, M, (H, Wc), generator=g).float().pin_memory()
    enc = crw.Resnet(pos_embed=False).cuda().eval()
    lp = crw.LabelPropVOS_CRW({"CXT_SIZE": LP["ctx"], "RADIUS": LP["radius"], "TEMP": LP["temp"], "KNN": LP["k"]})
    out = [None]
    t_enc = [0.0]

    def step(i):
        rg = rg_host.cuda(non_blocking=True)
        seg = seg_host.cuda(non_blocking=True)
        ds = crw.RGDataset(rg, length=Tl, dim=LP["patch"], overlap=LP["overlap"])
        seq = ds[0]                                                        # [T,N,h,w], one item = the whole radargram
        pred, _, _ = crw.propagate(seq, seg[:, :LP["patch"][1]], enc, lp, M, False, False)
        up = crw.ops.labels_upsample(pred.t().contiguous().to(torch.int32)[None], H, Wc)
        out[0] = up.cpu()                                                  # D2H of the label map

    ms = timed_loop(step, max(2, min(args.steps, 5)), 2, world, flush=flush)

    # the encoder's share, timed alone on the same frames
    ds = crw.RGDataset(rg_host.cuda(), length=Tl, dim=LP["patch"], overlap=LP["overlap"])
    seq = ds[0]
    x = seq.reshape(-1, LP["patch"][0], LP["patch"][1]).unsqueeze(1)

    def enc_step(i):
        with torch.no_grad():
            enc(x)

    ms_enc = timed_loop(enc_step, max(2, min(args.steps, 5)), 2, world)
    cols = Wc * world
    return dict(value=cols / (ms * 1e-3), unit="columns/s", ms_per_step=ms, encoder_ms=ms_enc,
                h2d_bytes_per_step=int(rg_host.numel() * 4 + seg_host.numel() * 4), d2h_bytes_per_step=int(H * Wc * 4),
                what="host radargram + segmentation -> RGDataset -> frames -> Resnet encoder (PyTorch, eval, fp32) -> propagate (exact tensor "
                     "path) -> nearest upsample -> label map on the host; reference call pattern scripts/test/test_all.py:91-96, one item = the "
                     "whole radargram; includes the horizontality metric's D2H and (no ruptures here) no change-point search")


def lp_call_traffic(cfg5, lp_precision):
    """Whole-call DRAM traffic of one label-propagation call: dram__bytes_read.sum + dram__bytes_write.sum summed over every kernel
    of the call, from the ncu capture summarised in profiles/r02_lp_call_dram.json (written by tools/summarize_dram.py from an
    `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` run of tools/lp_once.py).  None when no capture covers the case."""
    path = os.path.join(ROOT, "profiles", "r02_lp_call_dram.json")
    key = ("cfg5_" if cfg5 else "cfg3_") + ("fp32" if lp_precision == "fp32" else "tc_exact")
    try:
        d = json.load(open(path))[key]
        per = d["bytes"] / d.get("radargrams_in_capture", 1)
        return dict(bytes_per_radargram=int(per), source=f"profiles/r02_lp_call_dram.json[{key}] <- {d['capture']}")
    except Exception:  # noqa: BLE001
        return dict(bytes_per_radargram=None, source="no ncu capture for this case")


def run_b200(args):
    import torch.distributed as dist
    rank, world, local = dist_env()
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local, world)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import radar_sounder_crw_b200 as crw
    pk = peaks()
    torch.manual_seed(11 + rank)
    only = args.only
    tr = bench_train(crw, args, rank, world, local, pk) if only in ("all", "train") else None
    # the hot path alone.  Primary = what the train step runs: precision=BF16X3, which at this size takes the fused tcgen05 kernels
    # (walk_fused.cu, role-split: one launch per direction); beside it the fp32 shared-memory kernels, the eight-kernel BF16X3 engine
    # (CRW_WALK_FUSED=0), and both BF16X3 engines at a batch that fills the SMs (one CTA per batch element there)
    wk = wk_f32 = wk_tc = wk_fused = None
    if only in ("all", "walk"):
        fused_note = "walk_fused_fwd_roles_kernel + walk_fused_bwd_roles_kernel (tcgen05 kind::f16 on bf16 hi/lo pairs, frames by TMA, tiles by bulk copies), 2 launches"
        wk = bench_walk(crw, args, world, pk, precision=crw.ops.PREC_BF16X3, kernel_note=fused_note, launches=2)
        wk_f32 = bench_walk(crw, args, world, pk)
        os.environ["CRW_WALK_FUSED"] = "0"
        try:
            wk_tc = bench_walk(crw, args, world, pk, precision=crw.ops.PREC_BF16X3,
                               kernel_note="walk fwd+bwd kernels (BF16X3: mma.sync on bf16 hi/lo pairs), 8 launches")
            rd128 = bench_walk(crw, args, world, pk, B=4 * TRAIN["B"], precision=crw.ops.PREC_BF16X3)
        finally:
            del os.environ["CRW_WALK_FUSED"]
        rf128 = bench_walk(crw, args, world, pk, B=4 * TRAIN["B"], precision=crw.ops.PREC_BF16X3,
                           kernel_note="walk_fused_fwd_kernel + walk_fused_bwd_kernel (one CTA per batch element), 2 launches", launches=2)
        wk_fused = {f"B{TRAIN['B']}": dict(ms=wk["ms"], ms_eager_dispatch=wk["ms_eager"], launches=2, kernels="role-split (4 CTAs per batch element)",
                                          eight_kernel_engine_ms=wk_tc["ms"], eight_kernel_engine_ms_eager=wk_tc["ms_eager"]),
                    f"B{4 * TRAIN['B']}": dict(ms=rf128["ms"], ms_eager_dispatch=rf128["ms_eager"], launches=2, kernels="one CTA per batch element",
                                              tflops=rf128["roofline"]["achieved"], eight_kernel_engine_ms=rd128["ms"],
                                              eight_kernel_engine_ms_eager=rd128["ms_eager"])}
    # the large-N tile engine (walk_tc_tiles.cu) at the scaled geometries of SURVEY appendix D -- never reference-size figures
    wk_scaled = None
    if only == "all" and world == 1:
        wk_scaled = []
        for n in (185, 369):
            r = bench_walk(crw, args, world, pk, N=n, T=20, B=32, precision=crw.ops.PREC_BF16X3, launches=51,
                           kernel_note="persistent tcgen05 tile engine (TMA-fed bf16 hi/lo planes, 3 MMA passes; algorithmic FLOPs counted once)")
            wk_scaled.append(dict(shape=r["shape"], ms=r["ms"], launches=51, tflops_algorithmic=r["roofline"]["achieved"],
                                  frac_of_bf16_peak=r["roofline"]["frac"], note="scaled geometry (SURVEY appendix D), not a reference-size figure"))
        torch.cuda.empty_cache()
    lp = bench_labelprop(crw, args, rank, world, pk) if only in ("all", "labelprop") else None
    # the multi-GPU configurations BASELINE names (config 4: T=20 + UNet-as-encoder, data parallel; config 5: 64 radargrams of
    # 50k columns sharded over the ranks, strong scaling) ride in the same line as objects of their own
    tr4 = bench_train(crw, args, rank, world, local, pk, cfg4=True) if only in ("all", "train4") and not args.no_cfg45 else None
    lp5 = bench_labelprop(crw, args, rank, world, pk, lp_config=5) if only in ("all", "labelprop5") and not args.no_cfg45 else None
    if tr4 is not None:
        torch.cuda.empty_cache()
    if only == "walk_tc_large":     # profiling aid: the tcgen05 walk engine at the scaled geometry N=369
        r = bench_walk(crw, args, world, pk, N=369, T=20, B=32, precision=crw.ops.PREC_BF16X3,
                       kernel_note="walk fwd+bwd kernels, tcgen05 bf16x3 GEMMs, 8 launches")
        if rank == 0:
            emit(dict(only=only, walk=r))
        return
    if only == "walk_sweep":
        sweep = bench_walk_sweep(crw, args, world, pk)
        if rank == 0:
            emit(dict(only=only, walk_sweep=sweep))
        return

    h2d = h2d_rate(world) if only == "all" else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and only == "all":
        cpu = cpu_baselines(steps_train=1, lp_frames=lp["frames"])
        lp["cpu_baseline"] = cpu["labelprop"]
        if lp5 is not None:
            lp5["cpu_baseline"] = cpu["labelprop5"]

    if rank == 0:
        if only != "all":
            emit(dict(only=only, train=tr, walk=wk, labelprop=lp, train_cfg4=tr4, labelprop_cfg5=lp5))
        else:
            B, T = TRAIN["B"], TRAIN["T"]
            hot = dict(what="walk fwd+bwd (crw_b200::walk_loss + backward) as the train step runs it: precision=BF16X3 on the fused tcgen05 kernels "
                            "(one launch per direction, role-split CTAs handing operand tiles over through L2; frames by TMA; softmax / lse-diag / "
                            "softmax backward in the MMA epilogues; no N x N matrix through global memory between steps); embeddings resident, "
                            "CUDA-graph replay; loss 1e-7 / dx 1e-5 of the fp64 oracle (tests/test_gpu_walk_fused.py)",
                       ms=wk["ms"], ms_eager_dispatch=wk["ms_eager"],
                       launches=wk["launches"], share_of_step=wk["ms"] / tr["ms_per_step"],
                       fp32=dict(ms=wk_f32["ms"], ms_eager_dispatch=wk_f32["ms_eager"], launches=8,
                                 note="precision=FP32: shared-memory-resident fp32 FMA kernels (round 1's train-step path)"),
                       bf16x3_eight_kernels=dict(ms=wk_tc["ms"], ms_eager_dispatch=wk_tc["ms_eager"], launches=8,
                                                 note="precision=BF16X3 with CRW_WALK_FUSED=0: shared-memory kernels, warp-level mma.sync"),
                       fused_by_batch=wk_fused, scaled_geometry_tile_engine=wk_scaled)
            line = dict(
                metric="crw_train_radargrams_per_sec", value=tr["value"], unit="radargrams/s", n_gpus=world,
                steps=args.steps, warmup=args.warmup, ms_per_step=tr["ms_per_step"], higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="bf16x3 walk (tcgen05, fp32 accumulate); encoder fp32", data="synthetic",
                config=dict(workload="BASELINE config 2: CRW train step, B=32 per GPU, T=10 frames, 400-row radargram, "
                                     "32x32 patches overlap (24,0) -> N=47 nodes, ResNet[1,1,1,1] encoder (PyTorch fp32), "
                                     "fused tcgen05 walk fwd+bwd (bf16x3, two launches), Adam; random-init weights",
                            global_batch=B * world, frames=T, nodes=tr["N"], tau=TRAIN["tau"],
                            parallelism=f"dp{world}" + (" (one process per GPU, ONE NCCL all-reduce of the 19.9 MB of encoder gradients per step: parallel.FlatGradients)" if world > 1 else ""),
                            l2="4 rotating input batches (246 MB) > L2"),
                clocks=tr["clocks"], e2e=tr["e2e"], gpu_launches=(wk["launches"] + (1 if tr["optimizer"].startswith("optim.FlatAdam") else 0)) * args.steps,
                roofline=wk["roofline"], hot_path=hot, loss=tr["loss"],
                cpu_baseline=(cpu["train"] if cpu else None), labelprop=lp)
            line["host_link"] = dict(h2d_gbs_per_rank_all_ranks_copying=h2d, rank0_numa=numa,
                                     note="pinned host -> device, 256 MB x 4 per rank, every rank at once: the floor under every e2e figure")
            if cpu:
                hot["cpu_baseline"] = cpu["walk"]
            if tr4 is not None:
                line["train_cfg4"] = dict(
                    metric="crw_train_radargrams_per_sec", value=tr4["value"], unit="radargrams/s", ms_per_step=tr4["ms_per_step"],
                    steps=tr4["steps"], warmup=tr4["warmup"], scaling="weak", e2e=tr4["e2e"], loss=tr4["loss"], clocks=tr4["clocks"],
                    config=dict(workload="BASELINE config 4: data-parallel CRW train step, B=32 per GPU, T=20 frames, N=47 nodes, "
                                         "UNet-as-encoder adapter (UNet(1,128) + global average pool, 2048-patch chunks under activation "
                                         "checkpointing; NOT reference behaviour: the reference never feeds CRW from its UNet), fused tcgen05 "
                                         "walk fwd+bwd (bf16x3), Adam",
                                global_batch=B * world, frames=20, nodes=tr4["N"],
                                parallelism=f"dp{world}" + (" (one process per GPU, ONE NCCL all-reduce of the 17.2 MB of encoder gradients per step: parallel.FlatGradients)" if world > 1 else "")))
            if lp5 is not None:
                line["labelprop_cfg5"] = lp5
            emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle port) -- the only place bench.py touches oracle/
# ------------------------------------------------------------------------------------------------
def cpu_train_port(steps, warmup, B):
    from oracle.walk_torch_port import make_resnet_encoder, train_step
    torch.manual_seed(11)
    encoder = make_resnet_encoder().train()
    opt = torch.optim.Adam(encoder.parameters(), lr=TRAIN["lr"])
    seq = synth_train_batch(B, TRAIN["T"], 5)
    best = None
    for threads in sorted({1, os.cpu_count() or 1}):
        torch.set_num_threads(threads)
        for _ in range(warmup):
            train_step(encoder, opt, seq, TRAIN["tau"])
        t0 = time.perf_counter()
        for _ in range(steps):
            train_step(encoder, opt, seq, TRAIN["tau"])
        dt = (time.perf_counter() - t0) / steps
        log(f"cpu train port: B={B} threads={threads} {dt:.2f} s/step")
        if best is None or dt < best[0]:
            best = (dt, threads)
    return best


def cpu_baselines(steps_train, lp_frames):
    from oracle import c_oracle
    loaded = load_reference()
    if loaded is not None:
        dt, threads, every = reference_train_step(loaded[1], TRAIN_CPU_SAMPLE_B, steps_train, 1)
        train = dict(value=TRAIN_CPU_SAMPLE_B / dt, unit="radargrams/s", cores=threads, kind="reference",
                     sample=f"B={TRAIN_CPU_SAMPLE_B} of the B=32 batch (same T=10, N=47): the reference's own CRW + Resnet + Adam "
                            f"(baseline/_ref, scripts/train.py:56-72 verbatim) on the host cores; s/step by threads: {every}")
    else:
        dt, threads = cpu_train_port(steps_train, 1, TRAIN_CPU_SAMPLE_B)
        train = dict(value=TRAIN_CPU_SAMPLE_B / dt, unit="radargrams/s", cores=threads, kind="port",
                     sample=f"B={TRAIN_CPU_SAMPLE_B} of the B=32 batch (same T=10, N=47, ResNet encoder, reference-order "
                            f"(T-2)^2 walk with autograd, Adam), torch CPU, faster of 1 and {os.cpu_count()} threads")
    rs = np.random.RandomState(3)
    Nl = 49
    feats = rs.randn(1, lp_frames, Nl, LP["C"]).astype(np.float32)
    l0 = rs.randint(0, LP["M"], (1, Nl)).astype(np.int32)
    c_oracle.labelprop(feats[:, :64], l0, LP["M"], LP["ctx"], LP["radius"], LP["temp"], LP["k"], want_masks=False, want_topk=False)
    t0 = time.perf_counter()
    c_oracle.labelprop(feats, l0, LP["M"], LP["ctx"], LP["radius"], LP["temp"], LP["k"], want_masks=False, want_topk=False)
    dt_lp = time.perf_counter() - t0
    lp = dict(value=lp_frames * COLS_PER_FRAME / dt_lp, unit="columns/s", cores=c_oracle.num_threads(), kind="port",
              sample=f"full config-3 radargram ({lp_frames} frames), linear-time C port (oracle/crw_oracle.c, OpenMP); the "
                     "reference's own O(T^2) torch loop measured 264-325 columns/s on 8 cores (BASELINE.md)")
    # config 5: one of the 64 radargrams (400 x 50000 columns, k=20, r=24) on the C port, times 64
    T5 = LP5["cols"] // COLS_PER_FRAME
    feats5 = rs.randn(1, T5, Nl, LP5["C"]).astype(np.float32)
    t0 = time.perf_counter()
    c_oracle.labelprop(feats5, l0, LP5["M"], LP5["ctx"], LP5["radius"], LP5["temp"], LP5["k"], want_masks=False, want_topk=False)
    dt5 = time.perf_counter() - t0
    lp5 = dict(value=T5 * COLS_PER_FRAME / dt5, unit="columns/s", cores=c_oracle.num_threads(), kind="port",
               sample=f"ONE of the 64 radargrams of config 5 ({T5} frames, k=20, r=24) on the linear-time C port (OpenMP), {dt5:.2f} s; "
                      f"the 64 radargrams are independent, so the whole job is 64 x that ({64 * dt5:.1f} s) at the same columns/s")
    # the hot path alone on the host: the reference's own walk (src/model.py:22-46 + autograd) when baseline/_ref is there
    loaded = load_reference()
    if loaded is not None:
        wdt, wthreads, wevery = reference_walk_only(loaded[1], TRAIN["B"], TRAIN["T"], 47, 2, 1)
        walk = dict(value=TRAIN["B"] / wdt, unit="radargrams/s", ms_per_step=wdt * 1e3, cores=wthreads, kind="reference",
                    sample=f"walk only, reference code verbatim (baseline/_ref/src/model.py:22-46 + autograd, pre-computed embeddings), "
                           f"full B={TRAIN['B']}, T={TRAIN['T']}, N=47; s/step by threads: {wevery}")
    else:
        walk = None
    return dict(train=train, labelprop=lp, labelprop5=lp5, walk=walk)


REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def load_reference():
    """The UNMODIFIED reference (baseline/_ref/src, staged by __graft_entry__.build()) imported through oracle/ref_shim.py with the
    'cuda' literals mapped to the host: returns the shim namespace, or None when the copy is absent."""
    if not os.path.isfile(os.path.join(REF_DIR, "src", "model.py")):
        return None
    os.environ["CRW_REFERENCE_ROOT"] = REF_DIR
    os.environ["CRW_REFERENCE_FORCE_CPU"] = "1"
    from oracle import ref_shim
    return ref_shim, ref_shim.load()


class _Precomputed(torch.nn.Module):
    """Stands in for the encoder in the walk-only leg: returns pre-computed per-patch features."""

    def __init__(self, feats):
        super().__init__()
        self.feats = feats

    def forward(self, _x):
        return self.feats.reshape(-1, self.feats.shape[-1])


def _best_over_threads(fn, steps, warmup):
    """fn() timed at 1 thread and at all host threads (the hot path's many small ops are often faster on one): (s/step, threads, all)."""
    best, every = None, {}
    for threads in sorted({1, os.cpu_count() or 1}):
        torch.set_num_threads(threads)
        for _ in range(warmup):
            fn()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        dt = (time.perf_counter() - t0) / steps
        every[threads] = dt
        if best is None or dt < best[0]:
            best = (dt, threads)
    return best[0], best[1], every


def reference_train_step(ref, B, steps, warmup):
    """scripts/train.py:56-72 verbatim on the host cores: reference CRW + reference Resnet + Adam on a B-radargram sample."""
    import contextlib
    torch.manual_seed(11)
    with open(os.devnull, "w") as dn, contextlib.redirect_stdout(dn):
        enc = ref.encoder.Resnet(pos_embed=False)
    model = ref.model.CRW(enc, TRAIN["tau"], False)
    model.train(True)
    opt = torch.optim.Adam(model.parameters(), lr=TRAIN["lr"])
    seq = synth_train_batch(B, TRAIN["T"], 5)

    def step():
        loss, _ = model(seq)
        opt.zero_grad()
        loss.backward()
        opt.step()

    return _best_over_threads(step, steps, warmup)


def reference_walk_only(ref, B, T, N, steps, warmup):
    """src/model.py:22-46 + autograd with pre-computed embeddings (no encoder): the hot path alone, reference code verbatim."""
    torch.manual_seed(11)
    x = torch.randn(B, T, N, 128, requires_grad=True)
    model = ref.model.CRW(_Precomputed(x), TRAIN["tau"], False)
    dummy = torch.zeros(B, T, N, 2, 2)

    def step():
        x.grad = None
        loss, _ = model(dummy)
        loss.backward()

    return _best_over_threads(step, steps, warmup)


def reference_propagate(shim, ref, T, steps):
    """src/utils.py:94-161 verbatim (its own O(T^2) frame loop) on pre-computed features of T frames."""
    torch.manual_seed(11)
    N = 49
    feats = torch.randn(T, N, LP["C"])
    seg_ref = torch.randint(0, LP["M"], (400, 8))
    lp = ref.labelprop.LabelPropVOS_CRW({"CXT_SIZE": LP["ctx"], "RADIUS": LP["radius"], "TEMP": LP["temp"], "KNN": LP["k"]})

    def run():
        with shim.cpu_device_patches():
            ref.utils.propagate(torch.zeros(T, N, 2, 2), seg_ref, _Precomputed(feats), lp, LP["M"], False, False)

    return _best_over_threads(run, steps, 0)


def run_reference(args):
    """CPU arm: the reference's own code (baseline/_ref, kind "reference") on the box's host cores; the oracle ports only where the
    reference itself cannot finish in bounded time (label propagation at full config-3 length: the reference loop is O(T^2))."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from oracle import c_oracle
    B = TRAIN_CPU_SAMPLE_B
    loaded = load_reference()
    steps = max(1, min(args.steps, 3))
    if loaded is not None:
        shim, ref = loaded
        dt, threads, every = reference_train_step(ref, B, steps, 1)
        kind = "reference"
        wdt, wthreads, wevery = reference_walk_only(ref, TRAIN["B"], TRAIN["T"], 47, steps, 1)
        pdt, pthreads, pevery = reference_propagate(shim, ref, 300, 1)
        walk = dict(value=TRAIN["B"] / wdt, unit="radargrams/s", ms_per_step=wdt * 1e3, cores=wthreads, kind="reference",
                    sample=f"walk only (src/model.py:22-46 + autograd, pre-computed embeddings), full B={TRAIN['B']}, T={TRAIN['T']}, N=47; "
                           f"s/step by threads: {wevery}")
        prop = dict(value=300 * COLS_PER_FRAME / pdt, unit="columns/s", cores=pthreads, kind="reference",
                    sample=f"src/utils.py propagate verbatim (O(T^2) frame loop) on T=300 frames of pre-computed features; s by threads: {pevery}")
    else:
        dt, threads = cpu_train_port(steps, 1, B)
        every, kind, walk, prop = {threads: dt}, "port", None, None
    Tl = LP["cols"] // COLS_PER_FRAME
    rs = np.random.RandomState(3)
    feats = rs.randn(1, Tl, 49, LP["C"]).astype(np.float32)
    l0 = rs.randint(0, LP["M"], (1, 49)).astype(np.int32)
    t0 = time.perf_counter()
    for _ in range(steps):
        c_oracle.labelprop(feats, l0, LP["M"], LP["ctx"], LP["radius"], LP["temp"], LP["k"], want_masks=False, want_topk=False)
    dt_lp = (time.perf_counter() - t0) / steps
    value = B / dt
    sample = (f"each step = B={B} radargrams of the B=32 config-2 batch (T=10, N=47, reference Resnet encoder, reference CRW loss, "
              f"autograd, Adam: scripts/train.py:56-72 verbatim) on the host cores, {threads} threads; s/step by threads: {every}")
    line = dict(impl="reference", metric="crw_train_radargrams_per_sec", value=value, unit="radargrams/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=dt * 1e3, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="bf16x3 walk (tcgen05, fp32 accumulate); encoder fp32", data="synthetic",
                config=dict(workload="BASELINE config 2 (bounded CPU sample): " + sample),
                cpu_baseline=dict(value=value, unit="radargrams/s", cores=threads, kind=kind, sample=sample),
                e2e=dict(value=value, unit="radargrams/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                hot_path=walk,
                labelprop=dict(metric="labelprop_columns_per_sec", unit="columns/s", value=Tl * COLS_PER_FRAME / dt_lp,
                               ms_per_step=dt_lp * 1e3,
                               cpu_baseline=dict(value=Tl * COLS_PER_FRAME / dt_lp, unit="columns/s",
                                                 cores=c_oracle.num_threads(), kind="port",
                                                 sample="full config-3 radargram, linear-time C port (OpenMP)"),
                               reference_loop=prop,
                               e2e=dict(value=Tl * COLS_PER_FRAME / dt_lp, unit="columns/s", h2d_bytes_per_step=0,
                                        d2h_bytes_per_step=0)))
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lp-precision", default="both", choices=["both", "bf16x3", "fp32"])
    ap.add_argument("--lp-config", type=int, default=3, choices=[3, 5])
    ap.add_argument("--no-cfg45", action="store_true", help="skip the config-4 / config-5 objects (quick runs)")
    ap.add_argument("--only", default="all", choices=["all", "train", "walk", "labelprop", "walk_sweep", "walk_tc_large", "train4", "labelprop5"],
                    help="profiling aid: run one section only (the JSON line is then not the contract line)")
    args = ap.parse_args()
    protect_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_b200(args)


if __name__ == "__main__":
    main()
