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
    for n_obj in (5, 2, 1):
        q = ctx.Queue()
        port = 29600 + n_obj
        ps = [ctx.Process(target=_worker, args=(r, 2, port, n_obj, q)) for r in range(2)]
        [p.start() for p in ps]
        res = sorted(q.get(timeout=120) for _ in ps)
        [p.join(60) for p in ps]
        assert res == [(0, True, n_obj), (1, True, n_obj)]
