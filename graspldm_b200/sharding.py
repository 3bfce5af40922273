"""Multi-GPU generation: one process per GPU, objects sharded across ranks, no
collective on the compute path, ONE final all_gather of the results over NCCL (NVLink 5 / NVSwitch).
With gloo the same code runs on CPU tensors for the host-logic tests (the compute callback is injected)."""
import torch
import torch.distributed as dist


def shard_bounds(n_objects, world_size, rank):
    """Contiguous, balanced split: the first (n mod w) ranks get one extra object."""
    base, extra = divmod(n_objects, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_counts(n_objects, world_size):
    return [shard_bounds(n_objects, world_size, r)[1] - shard_bounds(n_objects, world_size, r)[0] for r in range(world_size)]


def gather_results(local, counts, group=None):
    """local: dict of tensors whose dim 0 is this rank's object count; returns the dict concatenated over ranks
    in rank order (every rank receives the full result).  Uneven shards are padded to the largest one."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    mx = max(counts)
    out = {}
    for k, v in local.items():
        pad = torch.zeros((mx,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
        pad[: v.shape[0]] = v
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        out[k] = torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
    return out


def generate_sharded(generate_fn, pcs, num_grasps, rank=None, world_size=None, gather=True, group=None):
    """generate_fn(local_pcs, first_object_index) -> dict of per-object tensors ([n_local, G, ...]).
    pcs: [n_objects, N, 3] (every rank holds, or can produce, the full list; only its slice is touched)."""
    rank = dist.get_rank(group) if rank is None else rank
    world_size = dist.get_world_size(group) if world_size is None else world_size
    lo, hi = shard_bounds(pcs.shape[0], world_size, rank)
    local = generate_fn(pcs[lo:hi], lo)
    if not gather or world_size == 1:
        return local
    return gather_results(local, shard_counts(pcs.shape[0], world_size), group)
