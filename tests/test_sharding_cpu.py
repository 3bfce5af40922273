"""CPU, world_size 2 over gloo: object sharding and the single final gather."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graspldm_b200 import sharding


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 2, 7, 8, 9, 64, 1025):
        for w in (1, 2, 4, 8):
            spans = [sharding.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sharding.shard_counts(n, w)


def _worker(rank, world, port, n_obj, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pcs = torch.arange(n_obj * 4 * 3, dtype=torch.float32).view(n_obj, 4, 3)

    def fake_generate(local, first):   # stands in for the CUDA path: per-object results depend only on the object
        return dict(grasp_tmrp=local.sum((1, 2)).view(-1, 1, 1).expand(-1, 2, 6).contiguous() + torch.arange(6.),
                    index=torch.arange(first, first + local.shape[0]).view(-1, 1))

    out = sharding.generate_sharded(fake_generate, pcs, 2)
    ref = fake_generate(pcs, 0)
    ok = all(torch.equal(out[k], ref[k]) for k in ref)
    q.put((rank, ok, int(out["index"].shape[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_generation_equals_single_rank():
    ctx = mp.get_context("spawn")
    for n_obj in (5, 2):          # (fewer objects than ranks: test_two_ranks_split_the_grasps_of_one_object)
        q = ctx.Queue()
        port = 29600 + n_obj
        ps = [ctx.Process(target=_worker, args=(r, 2, port, n_obj, q)) for r in range(2)]
        [p.start() for p in ps]
        res = sorted(q.get(timeout=120) for _ in ps)
        [p.join(60) for p in ps]
        assert res == [(0, True, n_obj), (1, True, n_obj)]


def test_plan_splits_grasps_when_objects_are_fewer_than_ranks():
    for n_obj, G, w in ((1, 20, 8), (3, 7, 8), (2, 5, 4), (7, 3, 8), (1, 3, 8)):
        seen = {}
        for r in range(w):
            lo, hi, g_lo, g_hi = sharding.plan(n_obj, G, w, r)
            assert hi == lo + 1 and 0 <= g_lo <= g_hi <= G
            seen.setdefault(lo, []).append((g_lo, g_hi))
        assert sorted(seen) == list(range(n_obj))
        for spans in seen.values():                       # the grasp ranges of an object tile [0, G) exactly once
            spans.sort()
            assert spans[0][0] == 0 and spans[-1][1] == G and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert sharding.plan(64, 20, 8, 3) == (24, 32, 0, 20)


def _worker_split(rank, world, port, n_obj, G, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pcs = torch.arange(n_obj * 4 * 3, dtype=torch.float32).view(n_obj, 4, 3)
    calls = []

    def fake_generate(local, first, g_n=G, g_first=0):    # results depend only on (object, grasp index)
        calls.append((int(first), int(local.shape[0]), g_n, g_first))
        gi = torch.arange(g_first, g_first + g_n, dtype=torch.float32)
        base = local.sum((1, 2)).view(-1, 1, 1) + 100.0 * gi.view(1, -1, 1)
        return dict(grasp_tmrp=(base + torch.arange(6.)).contiguous(), confidence=base[..., :1] * 0.5)

    out = sharding.generate_sharded(fake_generate, pcs, G)
    ref = fake_generate(pcs, 0)
    ok = all(out[k].shape == ref[k].shape and torch.equal(out[k], ref[k]) for k in ref)
    q.put((rank, ok, calls[0]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_split_the_grasps_of_one_object():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker_split, args=(r, 2, 29650, 1, 5, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert res == [(0, True, (0, 1, 3, 0)), (1, True, (0, 1, 2, 3))]
