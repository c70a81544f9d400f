"""The N>1 path on CPU: world_size-2 gloo processes shard bucket-pure batches with no data-path
collective, agree on max-over-ranks time and gather ragged per-sample metrics."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ragb_vae_b200 import sharding as S

    shapes = S.sample_bucket_batches(41, 8, seed=1234)
    owned = S.assign_batches(shapes, world)
    mine = owned[rank]
    # every rank computes the same partition; it is disjoint and complete
    flat = sorted(i for r in owned for i in r)
    assert flat == list(range(len(shapes)))
    # balanced within one batch of the mean
    loads = [sum(S.batch_cost(shapes[i]) for i in r) for r in owned]
    assert max(loads) - min(loads) <= max(S.batch_cost(s) for s in shapes) + 1e-9
    # job time = slowest rank
    t = S.max_over_ranks(10.0 + rank, torch.device("cpu"))
    assert t == 10.0 + (world - 1)
    # ragged per-sample gather: rank r contributes one value per owned batch
    metric = torch.tensor([float(i) for i in mine])
    allm = S.gather_per_sample(metric)
    assert sorted(allm.tolist()) == [float(i) for i in range(len(shapes))]
    # units processed by the whole job (the numerator of the weak-scaling metric)
    pix = torch.tensor([sum(b * h * w for (b, h, w) in (shapes[i] for i in mine))], dtype=torch.float64)
    dist.all_reduce(pix)
    assert pix.item() == sum(b * h * w for (b, h, w) in shapes)
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_world_size_2_gloo_sharding(tmp_path, lib_built):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_assignment_properties(lib_built):
    from ragb_vae_b200 import sharding as S

    shapes = S.sample_bucket_batches(256, 8)
    assert all(h % 32 == 0 and w % 32 == 0 and h * w <= 1048576 for _, h, w in shapes)
    frac_1024 = sum(1 for _, h, w in shapes if h == 1024 and w == 1024) / len(shapes)
    assert 0.7 < frac_1024 < 0.95
    for world in (1, 2, 4, 8):
        owned = S.assign_batches(shapes, world)
        assert sorted(i for r in owned for i in r) == list(range(256))
        loads = [sum(S.batch_cost(shapes[i]) for i in r) for r in owned]
        assert max(loads) / (sum(loads) / world) < 1.05
    with pytest.raises(ValueError):
        S.assign_batches(shapes, 0)


def _allreduce_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ragb_vae_b200.training import GradientAllReducer

    g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    red = GradientAllReducer(g, num_buckets=4)
    assert [hi - lo for lo, hi in red.buckets] == [250] * 4 and red.buckets[0] == (750, 1000)  # last parameters first
    # a boundary (decoder side begins at 600): no bucket straddles it, both sides are cut in proportion to their sizes
    red2 = GradientAllReducer(g, num_buckets=4, boundary=600)
    assert red2.buckets == [(800, 1000), (600, 800), (300, 600), (0, 300)]
    assert GradientAllReducer(g, num_buckets=4, boundary=950).buckets[0] == (950, 1000)
    # two boundaries (shallow encoder | deep encoder | decoder): every segment its own buckets, the spare one to the largest share
    red3 = GradientAllReducer(g, num_buckets=4, boundary=[100, 600])
    assert red3.buckets == [(600, 1000), (350, 600), (100, 350), (0, 100)]
    assert GradientAllReducer(g, num_buckets=2, boundary=[100, 600]).buckets == [(600, 1000), (0, 600)]
    for b in range(4):      # buckets become ready in backward order
        red.ready(b)
    scale = red.wait()
    assert scale == 1.0 / world
    assert torch.allclose(g * scale, torch.arange(1000, dtype=torch.float32) * (sum(range(1, world + 1)) / world))
    open(os.path.join(tmp, f"ar{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_world_size_2_gloo_gradient_allreduce(tmp_path, lib_built):
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_allreduce_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ar0").exists() and (tmp_path / "ar1").exists()


def test_trainer_bucket_schedule_covers_every_bucket_once(lib_built):
    """Host logic of VaeTrainStep's three-phase all-reduce schedule (no GPU): after the decoder's backward only buckets made
    of decoder / post_quant_conv parameters start; after the encoder's deep stages the buckets past the shallow segment; after
    the encoder's backward the rest; every bucket exactly once."""
    from types import SimpleNamespace

    from ragb_vae_b200.trainer import VaeTrainStep
    from ragb_vae_b200.training import GradientAllReducer

    names = ["encoder.a", "encoder.b", "decoder.a", "decoder.b", "decoder.c", "quant_conv.w", "post_quant_conv.w"]
    sizes = [300, 200, 250, 250, 300, 50, 50]
    offsets = [0]
    for n in sizes:
        offsets.append(offsets[-1] + n)
    step = object.__new__(VaeTrainStep)
    step.names = names
    step.opt = SimpleNamespace(offsets=offsets)
    started = []

    class FakeReducer(GradientAllReducer):
        def world(self):
            return 2

        def ready(self, bucket):
            started.append(bucket)

    step.reducer = FakeReducer(torch.zeros(offsets[-1]), num_buckets=7, boundary=[300, 550])
    step._shallow_end = 300   # "encoder.a" is the shallow segment
    step._mark_ready(0)
    first = list(started)
    for b in first:  # only decoder-side parameters inside these buckets
        lo, hi = step.reducer.buckets[b]
        inside = [n for i, n in enumerate(names) if offsets[i + 1] > lo and offsets[i] < hi]
        assert inside and all(n.startswith(("decoder.", "post_quant_conv.")) for n in inside), (b, inside)
    assert first, "some decoder-only bucket must be ready after the decoder's backward"
    step._mark_ready(1)
    second = [b for b in started if b not in first]
    assert second and all(step.reducer.buckets[b][0] >= 300 for b in second)   # deep encoder side only
    assert all(b in started for b, (lo, hi) in enumerate(step.reducer.buckets) if lo >= 300)
    step._mark_ready(2)
    assert sorted(started) == list(range(len(step.reducer.buckets)))
    assert [b for b in started if b not in first and b not in second] == [b for b, (lo, hi) in enumerate(step.reducer.buckets) if hi <= 300]
    # a bucket that mixes encoder-side and decoder-side parameters waits for the second phase
    mixed = [b for b, (lo, hi) in enumerate(step.reducer.buckets)
             if any(offsets[i + 1] > lo and offsets[i] < hi and not n.startswith(("decoder.", "post_quant_conv."))
                    for i, n in enumerate(names))]
    assert all(b not in first for b in mixed)


@pytest.mark.parametrize("arch", ["qwen", "flux"])
def test_trainer_encoder_split_is_a_contiguous_shallow_prefix(lib_built, arch):
    """Host logic behind the three-phase all-reduce (no GPU): the shallow phase of the encoder's backward is the stem plus the
    first blocks, their parameters are exactly the head [0, shallow_end) of the flat buffer and at most 10 % of the encoder side."""
    import ragb_vae_b200 as R
    from ragb_vae_b200.trainer import VaeTrainStep

    vae = R.RgbaAutoencoder(arch)
    step = object.__new__(VaeTrainStep)
    step.vae, step.flux = vae, arch == "flux"
    named = [(n, p) for n, p in vae.named_parameters() if ".time_conv." not in n]
    named = [np for np in named if not step._decoder_side(np[0])] + [np for np in named if step._decoder_side(np[0])]
    boundary = sum(p.numel() for n, p in named if not step._decoder_side(n))
    split, shallow_end = step._encoder_split(named, boundary)
    assert split >= 1 and shallow_end is not None and 0 < shallow_end <= 0.10 * boundary
    mods = [vae.encoder.conv_in] + [m for _, m in step._encoder_items()][:split]
    ids = {id(p) for m in mods for p in m.parameters()}
    off = 0
    for _, p in named:                      # the shallow modules' parameters are the first ones, nothing else in between
        if off < shallow_end:
            assert id(p) in ids
        else:
            assert id(p) not in ids
        off += p.numel()
