"""CPU, world_size 2 over gloo: the data-parallel path (scene sharding + one flat-gradient all-reduce) gives the
same update as a single process on the whole batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    torch.manual_seed(7)
    from latentsplat_b200.model.discriminator import DiscriminatorPatchGanCfg, get_discriminator
    m = get_discriminator(DiscriminatorPatchGanCfg("patch_gan", "kl_f8", base_dim=8, n_layers=2, pretrained=False), 3)
    for mod in m.modules():                                   # BatchNorm statistics are per-rank in the reference too;
        if isinstance(mod, torch.nn.BatchNorm2d):             # use eval-mode stats so that sharding is exactly linear
            mod.eval()
    return m


def _batch():
    g = torch.Generator().manual_seed(3)
    return {"image": torch.rand(4, 3, 32, 32, generator=g), "target": torch.rand(4, 1, 6, 6, generator=g)}


def _loss(m, b):
    return ((m(b["image"]) - b["target"]) ** 2).mean()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from latentsplat_b200.parallel import FlatGradients, shard_batch
    m = _model()
    fg = FlatGradients(m.parameters())
    fg.zero()
    _loss(m, shard_batch(_batch(), rank, world)).backward()
    fg.all_reduce_mean()
    if rank == 0:
        torch.save(fg.flat.clone(), out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_flat_gradient_allreduce_matches_single_process(tmp_path):
    from latentsplat_b200.parallel import FlatGradients, shard_batch
    out = str(tmp_path / "flat.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    m = _model()
    fg = FlatGradients(m.parameters())
    fg.zero()
    _loss(m, _batch()).backward()                 # mean over 4 scenes == mean of the two ranks' means over 2 scenes
    torch.testing.assert_close(got, fg.flat, rtol=1e-5, atol=1e-7)
    assert got.abs().max() > 0
    # grads are views of the flat buffer; zero() clears them; clipping scales the whole set
    assert all(p.grad.data_ptr() >= fg.flat.data_ptr() for p in fg.params)
    n0 = float(fg.norm())
    fg.clip_(0.5 * n0)
    assert float(fg.norm()) == pytest.approx(0.5 * n0, rel=1e-4)
    fg.zero()
    assert all(float(p.grad.abs().max()) == 0 for p in fg.params)
    sb = shard_batch({"a": torch.arange(8), "n": {"b": torch.arange(8)}}, 1, 2)
    assert sb["a"].tolist() == [1, 3, 5, 7] and sb["n"]["b"].tolist() == [1, 3, 5, 7]


def _worker_bucketed(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from latentsplat_b200.parallel import BucketedAllReduce, FlatGradients, shard_batch
    m = _model()
    fg = FlatGradients(m.parameters())
    red = BucketedAllReduce(fg, bucket_bytes=4096)            # several buckets for this small network
    order = []
    send = red._send
    red._send = lambda i: (order.append(i), send(i))[1]
    results = []
    for step in range(2):                                     # the hooks re-arm every step
        fg.zero()
        red.begin()
        _loss(m, shard_batch(_batch(), rank, world)).backward()
        sent_during_backward = len(order)
        red.finish()
        results.append(fg.flat.clone())
    if rank == 0:
        torch.save({"flat": results, "order": order, "n_buckets": len(red.buckets), "early": sent_during_backward}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_bucketed_overlapped_allreduce_matches_single_process(tmp_path):
    """BucketedAllReduce (hooks fire per bucket during backward, last layers first) == one all-reduce at the end."""
    from latentsplat_b200.parallel import FlatGradients
    out = str(tmp_path / "bucketed.pt")
    mp.spawn(_worker_bucketed, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    m = _model()
    fg = FlatGradients(m.parameters())
    fg.zero()
    _loss(m, _batch()).backward()
    for flat in got["flat"]:
        torch.testing.assert_close(flat, fg.flat, rtol=1e-5, atol=1e-7)
    assert got["n_buckets"] >= 3
    per_step = got["order"][: len(got["order"]) // 2]
    assert sorted(per_step) == list(range(got["n_buckets"])), "every bucket is reduced exactly once per step"
    assert per_step[0] == got["n_buckets"] - 1, "the bucket of the last layers goes first (backward order)"
