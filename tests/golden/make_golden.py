"""Generate the committed golden fixtures from the LIVE reference.

Run in the build container (the only place /root/reference exists):

    python tests/golden/make_golden.py

Nothing of the reference is copied: its modules are imported in place through
``oracle/ref_shim.py`` and *executed* on seeded synthetic inputs; only inputs and
outputs are saved (inputs are rounded to fp16-representable values so they store
exactly in half the bytes).

Fixtures
--------
walk_small_f64.npz   B=2,T=6,N=9,C=16, tau=0.07, fp64: x, loss, A, dx (autograd)
walk_t3_f64.npz      T=3 edge case (only k=1)
walk_cfg1_f32.npz    B=1,T=10,N=47,C=128, tau=0.07, fp32 (config-1 geometry)
walk_tau001_f32.npz  B=2,T=8,N=24,C=128, tau=0.01 (the reference's train default, train.py:31)
lp_quirk.npz         T=30,N=49,C=128,M=4, ctx=5,k=10,r=12: context-trim quirk active for n>6
lp_cfg3_short.npz    T=40,N=49,C=128,M=4, ctx=20,k=10,r=12 (config-3 parameters)
lp_masked_ties.npz   T=8,N=25,C=16,M=3, ctx=20,k=20,r=10 (test_all.py defaults: masked ids enter top-k)
lp_clustered.npz     T=26,N=47,C=128,M=4, ctx=20,k=10,r=12, clustered (near-collinear) features
io_unfold.npz        RGDataset items (dataset.py:19-47) at three geometries incl. both overlaps and get_smaller_item
io_seed.npz          Resize((N,1), NEAREST) + one-hot (utils.py:139-147) for (rows,N) pairs incl. non-divisible ones
io_fuse.npz          reversed-pass fusion (test_all.py:146-158), rules 0 / 1 / 3
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402


class _FixedEncoder(torch.nn.Module):
    """Stands in for the encoder: returns pre-computed per-patch features."""

    def __init__(self, feats):
        super().__init__()
        self.feats = feats

    def forward(self, _x):
        return self.feats.reshape(-1, self.feats.shape[-1])


def _round_fp16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.float16).to(t.dtype)


def make_walk(ref, name, B, T, N, C, tau, dtype, seed, clustered=False):
    torch.manual_seed(seed)
    x = torch.randn(B, T, N, C, dtype=dtype)
    if clustered:
        x = x + 3.0 * torch.randn(B, 1, 1, C, dtype=dtype)
    x = _round_fp16(x).requires_grad_(True)
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)  # model.py:36 builds the identity in the default dtype
    try:
        crw = ref.model.CRW(_FixedEncoder(x), tau, False)
        loss, A = crw(torch.zeros(B, T, N, 2, 2, dtype=dtype))
        if T >= 3:
            loss.backward()
            dx = x.grad.detach().numpy()
            loss_v = float(loss.item())
        else:
            dx = np.zeros(x.shape, x.detach().numpy().dtype)
            loss_v = float(loss)
    finally:
        torch.set_default_dtype(old)
    np.savez_compressed(os.path.join(HERE, name), x=x.detach().numpy().astype(np.float16), tau=tau,
                        loss=np.float64(loss_v), A=A.detach().numpy(), dx=dx,
                        dtype=str(dtype).replace("torch.", ""))
    print(name, "loss", loss_v)


def make_lp(ref, name, T, N, C, M, ctx, k, radius, temp, seed, clustered=False):
    torch.manual_seed(seed)
    feats = torch.randn(T, N, C)
    if clustered:
        feats = feats + 3.0 * torch.randn(1, 1, C)
    feats = _round_fp16(feats)
    seg_ref = torch.randint(0, M, (400, 8))
    lp = ref.labelprop.LabelPropVOS_CRW({"CXT_SIZE": ctx, "RADIUS": radius, "TEMP": temp, "KNN": k})
    masks = []
    orig_predict = lp.predict

    def predict(*a, **kw):
        m = orig_predict(*a, **kw)
        masks.append(m[0, :, :, 0].clone())
        return m

    lp.predict = predict
    with ref_shim.cpu_device_patches(), ref_shim.TopkSpy(ref) as spy:
        pred, xent, _ = ref.utils.propagate(torch.zeros(T, N, 2, 2), seg_ref, _FixedEncoder(feats), lp, M,
                                            False, False)
    np.savez_compressed(
        os.path.join(HERE, name), feats=feats.numpy().astype(np.float16), seg_col0=seg_ref[:, 0].numpy().astype(np.int16),
        M=M, ctx=ctx, k=k, radius=radius, temp=temp,
        labels=pred.numpy().astype(np.int16),                       # [N,T]
        W=torch.stack(spy.W).numpy(), I=torch.stack(spy.I).numpy().astype(np.int32),   # [T-1,k,N]
        masks=torch.stack(masks).numpy(),                           # [T-1,M,N]
        xent=xent.numpy())                                          # [N,T-1]
    print(name, "labels hist", np.bincount(pred.numpy().astype(np.int64).ravel(), minlength=M))


def make_io_unfold(ref, name):
    """Runs the reference RGDataset unmodified; only ``torch.load`` is pointed at a seeded tensor."""
    import contextlib
    out = {}
    cases = [("a", 40, 300, 5, (16, 16), (8, 0), False), ("b", 50, 211, 4, (32, 32), (24, 0), True),
             ("c", 37, 190, 3, (12, 10), (4, 6), False)]
    for tag, H, W, length, dim, overlap, flip in cases:
        torch.manual_seed(len(tag) + H)
        rg = _round_fp16(torch.randn(H, W))
        orig = torch.load
        ref.dataset.torch.load = lambda *_a, **_k: rg
        try:
            with open(os.devnull, "w") as dn, contextlib.redirect_stdout(dn):
                ds = ref.dataset.RGDataset(filepath="synthetic.pt", length=length, dim=dim, overlap=overlap, flip=flip)
        finally:
            ref.dataset.torch.load = orig
        idx = [0, len(ds) // 2, len(ds) - 1]
        out.update({f"{tag}_rg": rg.numpy().astype(np.float16), f"{tag}_geom": np.array([length, *dim, *overlap, int(flip)]),
                    f"{tag}_len": len(ds), f"{tag}_idx": np.array(idx),
                    f"{tag}_items": torch.stack([ds[i] for i in idx]).numpy().astype(np.float16),
                    f"{tag}_small": ds.get_smaller_item(1, 2).numpy().astype(np.float16)})
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "ok")


def make_io_seed(name):
    """utils.py:139-147 executed with the same torchvision transform the reference builds."""
    from torchvision import transforms
    from torchvision.transforms import InterpolationMode
    out = {}
    for tag, rows, N, M, W in [("a", 400, 49, 4, 16), ("b", 376, 47, 6, 32), ("c", 410, 49, 5, 16), ("d", 100, 33, 3, 8)]:
        torch.manual_seed(rows + N)
        seg = torch.randint(0, M, (rows, 3 * W)).float()
        labels, masks = [], []
        for col in (0, W, 2 * W):
            seg_ref = seg[:rows, col:col + W]
            down = transforms.Resize((N, 1), interpolation=InterpolationMode.NEAREST)
            label = down(seg_ref.unsqueeze(0)).squeeze(0)
            mask = torch.zeros(M, N, 1)
            for class_idx in range(0, M):
                mask[class_idx, :, :] = (label == class_idx).unsqueeze(0).float()
            labels.append(label.squeeze(1))
            masks.append(mask.squeeze(-1))
        out.update({f"{tag}_seg": seg.numpy().astype(np.int8), f"{tag}_geom": np.array([rows, N, M, W]),
                    f"{tag}_label0": torch.stack(labels).numpy().astype(np.int8), f"{tag}_mask0": torch.stack(masks).numpy()})
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "ok")


def make_io_fuse(name):
    """test_all.py:146-158 re-executed statement by statement (the script's main() is not importable)."""
    out = {}
    for rule, H, tot, rg_len in [(0, 24, 3, 40), (1, 30, 4, 25), (3, 17, 2, 36)]:
        torch.manual_seed(rule + H)
        final_pred = torch.randint(0, 6, (H, tot * rg_len)).float()
        rev_cat = torch.randint(0, 6, (H, tot * rg_len)).float()             # cat(seg_list, dim=1) of the reversed pass
        rev_cat[:, ::3][rev_cat[:, ::3] == 4] = 1                            # leave some columns free of class 4
        pred_seg_rev = rev_cat.unfold(dimension=1, size=rg_len, step=rg_len)
        pred_seg_rev = torch.flip(pred_seg_rev, (-1,)).view(pred_seg_rev.shape[0], -1)
        fp = final_pred.clone().flatten()
        if rule == 0:
            mask = pred_seg_rev.flatten() == 2
        if rule == 1:
            mask = torch.logical_and(pred_seg_rev.flatten() == 2, fp != 3)
            mask2 = torch.all(pred_seg_rev != 4, axis=0).unsqueeze(0).repeat([pred_seg_rev.shape[0], 1]).flatten()
            mask = torch.logical_and(mask, mask2)
        if rule == 3:
            mask = pred_seg_rev.flatten() == 2
            mask[:len(mask) // 2] = 0
        fp[mask] = 2
        out.update({f"r{rule}_fwd": final_pred.numpy().astype(np.int8), f"r{rule}_rev": rev_cat.numpy().astype(np.int8),
                    f"r{rule}_rg_len": rg_len, f"r{rule}_out": fp.view(H, -1).numpy().astype(np.int8)})
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "ok")


def main():
    ref = ref_shim.load()
    if "--io-only" in sys.argv:
        make_io_unfold(ref, "io_unfold.npz")
        make_io_seed("io_seed.npz")
        make_io_fuse("io_fuse.npz")
        return
    if "--round1b" in sys.argv:      # added later in round 1: BASELINE config 4 / config 5 parameters (the other fixtures stay byte-identical)
        make_walk(ref, "walk_cfg4_f32.npz", 1, 20, 47, 128, 0.07, torch.float32, 15)
        make_lp(ref, "lp_cfg5_short.npz", 45, 49, 128, 4, 20, 20, 24, 0.07, 16)
        return
    make_walk(ref, "walk_small_f64.npz", 2, 6, 9, 16, 0.07, torch.float64, 11)
    make_walk(ref, "walk_t3_f64.npz", 2, 3, 7, 8, 0.07, torch.float64, 12)
    make_walk(ref, "walk_cfg1_f32.npz", 1, 10, 47, 128, 0.07, torch.float32, 11)
    make_walk(ref, "walk_tau001_f32.npz", 2, 8, 24, 128, 0.01, torch.float32, 13, clustered=True)
    make_lp(ref, "lp_quirk.npz", 30, 49, 128, 4, 5, 10, 12, 0.07, 11)
    make_lp(ref, "lp_cfg3_short.npz", 40, 49, 128, 4, 20, 10, 12, 0.07, 12)
    make_lp(ref, "lp_masked_ties.npz", 8, 25, 16, 3, 20, 20, 10, 0.07, 13)
    make_lp(ref, "lp_clustered.npz", 26, 47, 128, 4, 20, 10, 12, 0.07, 14, clustered=True)
    make_io_unfold(ref, "io_unfold.npz")
    make_io_seed("io_seed.npz")
    make_io_fuse("io_fuse.npz")


if __name__ == "__main__":
    main()
