"""CPU, world_size 2, gloo: the data-parallel exchange of the hot path (GradReducer = the DDP replacement,
reference trainer.py:122-129,2257-2260): after the exchange every rank holds the MEAN of the per-rank flat gradients,
bucketed ranges included, and N-rank training equals single-rank training on the concatenated batch for a mean loss."""
import os
import socket
from types import SimpleNamespace

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unlearn_ft_b200.pdm.training.trainer import GradReducer
    g = torch.Generator().manual_seed(100 + rank)
    grad = torch.randn(1000, generator=g)
    model = SimpleNamespace(arena=SimpleNamespace(grad=grad.clone(), numel=1000))
    red = GradReducer(model)
    assert red.world == world
    red.reduce_range(600, 1000)       # bucketed, as issued per U-Net block from its backward (last block first)
    red.reduce_range(100, 300)
    red.reduce_all()                  # the complement: [0,100) and [300,600), each exactly once
    # the optimizer walks the buckets as their collectives land (FusedAdamW.step(ranges=...)): the slices it is handed
    # partition the arena exactly once, and the walk leaves the reducer ready for the next step
    spans = sorted(red.landed_ranges())
    assert spans[0][0] == 0 and spans[-1][1] == 1000 and all(a[1] == b[0] for a, b in zip(spans, spans[1:])), spans
    assert not red.pending and not red.done
    out[rank] = model.arena.grad.clone()
    dist.destroy_process_group()


def _worker_fused(rank, world, port, out):
    """Optimizer-inside-backward path: buckets handed over during 'backward' are exchanged and immediately given to the
    consumer; what no bucket claimed is exchanged by reduce_all() and left to landed_ranges()."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unlearn_ft_b200.pdm.training.trainer import GradReducer
    g = torch.Generator().manual_seed(200 + rank)
    grad = torch.randn(1000, generator=g)
    model = SimpleNamespace(arena=SimpleNamespace(grad=grad.clone(), numel=1000))
    red = GradReducer(model)
    seen = []                          # (lo, hi, copy of the slice at hand-over time) -- must already be the cross-rank mean

    def consumer(lo, hi):
        seen.append((lo, hi, model.arena.grad[lo:hi].clone()))
        model.arena.grad[lo:hi].zero_()            # what AdamW does (zero_grad fused into the update)

    red.on_landed = consumer
    red.reduce_range(700, 1000)
    red.reduce_range(200, 500)
    red.on_landed = None
    red.reduce_all()
    rest = list(red.landed_ranges())
    assert sorted(rest) == [(0, 200), (500, 700)], rest
    assert not red.pending and not red.done and not red.consumed
    out[rank] = dict(seen=[(lo, hi, t) for lo, hi, t in seen], grad=model.arena.grad.clone())
    dist.destroy_process_group()


def test_optimizer_in_backward_consumer_world2_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_fused, args=(world, port, out), nprocs=world, join=True)
    expect = sum(torch.randn(1000, generator=torch.Generator().manual_seed(200 + r)) for r in range(world)) / world
    for r in range(world):
        seen = out[r]["seen"]
        assert [(lo, hi) for lo, hi, _ in seen] == [(700, 1000), (200, 500)]
        for lo, hi, t in seen:
            assert torch.allclose(t, expect[lo:hi], atol=1e-6)           # the consumer saw the averaged gradient
        gr = out[r]["grad"]
        assert float(gr[700:1000].abs().sum()) == 0.0 and float(gr[200:500].abs().sum()) == 0.0
        assert torch.allclose(gr[0:200], expect[0:200], atol=1e-6) and torch.allclose(gr[500:700], expect[500:700], atol=1e-6)


def test_gradreducer_world2_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    expect = sum(torch.randn(1000, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)) / world
    for r in range(world):
        assert torch.allclose(out[r], expect, atol=1e-6)
    assert torch.equal(out[0], out[1])


def test_world1_is_noop():
    from unlearn_ft_b200.pdm.training.trainer import GradReducer
    g = torch.randn(10)
    red = GradReducer(SimpleNamespace(arena=SimpleNamespace(grad=g.clone(), numel=10)))
    red.reduce_all(), red.wait()
    assert torch.equal(red.arena.grad, g)


def test_optimizer_shards_partition_every_bucket():
    """Sharded optimizer (GradReducer._sharded_update): the owner slices of a bucket are disjoint, 256-byte aligned relative to
    the bucket start, cover its first m elements exactly once over the ranks, and leave fewer than world * 64 elements to the
    replicated remainder -- for every world size the bench runs and ragged bucket lengths."""
    from unlearn_ft_b200.pdm.training.trainer import GradReducer
    for world in (2, 4, 8):
        for start, end in ((0, 1000), (128, 128 + 64 * world), (4096, 4096 + 10_000_019), (64, 64 + world * 64 - 1)):
            spans = [GradReducer.shard_span(start, end, world, r) for r in range(world)]
            m, s = spans[0][0], spans[0][1]
            assert all(sp[0] == m and sp[1] == s for sp in spans)
            assert m % (world * 64) == 0 and s * world == m and 0 <= (end - start) - m < world * 64
            assert [sp[2] for sp in spans] == [start + r * s for r in range(world)]
            assert all((sp[2] - start) % 64 == 0 for sp in spans)


def test_block_buckets_partition_the_arena():
    """Per-block buckets (fired from each block's backward) are disjoint, contiguous arena slices; together with the
    complement handled by reduce_all() every gradient element is exchanged exactly once."""
    from unlearn_ft_b200.pdm.models import HyperStructure, UNet2DConditionModelPruned
    from unlearn_ft_b200.pdm.models.unet.unet_2d_conditional import SD21_CONFIG, structure_from_config
    from unlearn_ft_b200.pdm.training.trainer import GradReducer
    torch.manual_seed(0)
    av = HyperStructure.get_random_arch_vector(0.55, structure_from_config(SD21_CONFIG))
    model = UNet2DConditionModelPruned(arch_vector=av, device="meta", seed=None)
    buckets = GradReducer.block_buckets(model)
    names = {attr for (_, attr) in buckets}
    assert names == {"_grad_ready", "_grad_ready_head", "_grad_ready_stem"}
    assert len(buckets) == 4 + 1 + 4 + 2                      # down x4, mid, up x4, head, stem
    spans = sorted(buckets.values())
    for (l0, h0), (l1, h1) in zip(spans, spans[1:]):
        assert h0 <= l1                                        # disjoint
    covered = sum(h - l for l, h in spans)
    assert covered >= 0.999 * model.arena.numel                # (alignment padding aside) the blocks own everything
    for (mod, attr), (lo, hi) in buckets.items():              # every parameter of a block lies inside its bucket
        if attr != "_grad_ready":
            continue
        for sub in mod.modules():
            for a, (o, n) in getattr(sub, "_arena_off", {}).items():
                assert lo <= o and o + n <= hi
