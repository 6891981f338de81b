"""World-size-2 tests of the multi-GPU plumbing on CPU (gloo): sharding, gradient all-reduce, label gather.

The CUDA ops have no CPU path, so the per-rank compute here is the torch port of the walk (oracle); what is under
test is the host-side logic that bench.py / training scripts use around the ops.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _enc():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(16, 8))


def _worker(rank, world, port, ret, B=4, weighted=False, flat=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.walk_torch_port import crw_loss_reference_order
    from radar_sounder_crw_b200.parallel import allreduce_gradients, gather_labels, shard_range
    torch.manual_seed(11)
    T, N = 5, 6
    seq = torch.randn(B, T, N, 4, 4)
    b0, b1 = shard_range(B, rank, world)
    enc = _enc()
    # a parameter that only rank 0 uses: on rank 1 its .grad stays None and must still take part in the flat all-reduce
    extra = torch.nn.Parameter(torch.ones(8))
    emb = enc(seq[b0:b1].reshape(-1, 1, 4, 4)).reshape(b1 - b0, T, N, -1)
    if rank == 0:
        emb = emb * extra
    loss, _ = crw_loss_reference_order(emb, 0.07)
    if flat:
        from radar_sounder_crw_b200.parallel import FlatGradients
        fg = FlatGradients(list(enc.parameters()) + [extra], world)
        fg.zero()
        loss.backward()
        fg.reduce()
        views_kept = all(p.grad.untyped_storage().data_ptr() == fg.flat.untyped_storage().data_ptr() for p in enc.parameters())
        if rank == 0:
            ret["views_kept"] = views_kept
    else:
        loss.backward()
    if flat:
        pass
    elif weighted:
        allreduce_gradients(list(enc.parameters()) + [extra], world, local_items=b1 - b0, total_items=B)
    else:
        allreduce_gradients(list(enc.parameters()) + [extra], world)
    assert extra.grad is not None
    labels = torch.arange(b0, b1).view(-1, 1).repeat(1, 3)
    allv = gather_labels(labels, B, rank, world)
    if rank == 0:
        ret["grad"] = [p.grad.clone() for p in enc.parameters()]
        ret["labels"] = allv.clone()
    dist.destroy_process_group()


def test_shard_range_partitions():
    from radar_sounder_crw_b200.parallel import shard_range
    for n in [0, 1, 7, 8, 64, 65]:
        for w in [1, 2, 3, 8]:
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_two_rank_gradient_allreduce_equals_full_batch():
    from oracle.walk_torch_port import crw_loss_reference_order
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    torch.manual_seed(11)
    B, T, N = 4, 5, 6
    seq = torch.randn(B, T, N, 4, 4)
    enc = _enc()
    loss, _ = crw_loss_reference_order(enc(seq.reshape(-1, 1, 4, 4)).reshape(B, T, N, -1), 0.07)
    loss.backward()
    for g, p in zip(ret["grad"], enc.parameters()):
        assert torch.allclose(g, p.grad, rtol=1e-5, atol=1e-8)
    assert np.array_equal(ret["labels"].numpy(), np.arange(B)[:, None].repeat(3, 1))


def test_two_rank_uneven_shards_weighted_allreduce_equals_full_batch():
    """B = 5 on two ranks (3 + 2 items): weighting every rank's gradient by local/total gives the full-batch mean-loss gradient."""
    from oracle.walk_torch_port import crw_loss_reference_order
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret, 5, True), nprocs=2, join=True)
    torch.manual_seed(11)
    B, T, N = 5, 5, 6
    seq = torch.randn(B, T, N, 4, 4)
    enc = _enc()
    extra = torch.ones(8)
    emb = enc(seq.reshape(-1, 1, 4, 4)).reshape(B, T, N, -1)
    emb = torch.cat([emb[:3] * extra, emb[3:]])              # rank 0 (items 0..2) multiplies by the extra parameter (= 1)
    # the loss is a mean over batch elements of per-element terms, so the full-batch gradient is the item-weighted mean
    loss, _ = crw_loss_reference_order(emb, 0.07)
    loss.backward()
    for g, p in zip(ret["grad"], enc.parameters()):
        assert torch.allclose(g, p.grad, rtol=1e-5, atol=1e-8)


def test_two_rank_flat_gradients_equal_full_batch():
    """FlatGradients: the parameters' .grad are views of one buffer, autograd accumulates into them in place, ONE all-reduce."""
    from oracle.walk_torch_port import crw_loss_reference_order
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret, 4, False, True), nprocs=2, join=True)
    assert ret["views_kept"]
    torch.manual_seed(11)
    B, T, N = 4, 5, 6
    seq = torch.randn(B, T, N, 4, 4)
    enc = _enc()
    loss, _ = crw_loss_reference_order(enc(seq.reshape(-1, 1, 4, 4)).reshape(B, T, N, -1), 0.07)
    loss.backward()
    for g, p in zip(ret["grad"], enc.parameters()):
        assert torch.allclose(g, p.grad, rtol=1e-5, atol=1e-8)
