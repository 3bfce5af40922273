"""Multi-GPU generation: one process per GPU, no collective on the compute path, ONE final
all_gather of the packed results over NCCL (NVLink 5 / NVSwitch).

  * objects >= ranks: rank r takes a contiguous, balanced slice of the objects and all their grasps (the object latent
    stays local; reference semantics `z_pc.repeat_interleave(num_grasps)`, R/grasp_ldm/models/grasp_ldm.py:207);
  * objects <  ranks: the ranks are split into one group per object and the grasps of the object are split over the
    group; every rank of a group encodes the object itself (8 GFLOP - cheaper than broadcasting the 768-byte latent
    behind a collective on the critical path).

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


def plan(n_objects, num_grasps, world_size, rank):
    """-> (obj_lo, obj_hi, g_lo, g_hi): the objects and the grasp range (of each of them) this rank generates."""
    if n_objects >= world_size or n_objects == 0:
        lo, hi = shard_bounds(n_objects, world_size, rank)
        return lo, hi, 0, num_grasps
    for o in range(n_objects):                       # one group of ranks per object, group sizes differ by at most 1
        r_lo, r_hi = shard_bounds(world_size, n_objects, o)
        if r_lo <= rank < r_hi:
            g_lo, g_hi = shard_bounds(num_grasps, r_hi - r_lo, rank - r_lo)
            return o, o + 1, g_lo, g_hi
    raise AssertionError("unreachable")


def _pack(tensors, lead):
    """list of tensors sharing their first `lead` dims -> one buffer [..lead dims.., F] (dtype of the first)."""
    head = tuple(tensors[0].shape[:lead])
    return torch.cat([t.reshape(head + (-1,)).to(tensors[0].dtype) for t in tensors], dim=lead)


def _unpack(buf, keys, shapes, lead):
    out, off = {}, 0
    head = tuple(buf.shape[:lead])
    for k, shape in zip(keys, shapes):
        w = 1
        for s in shape:
            w *= s
        out[k] = buf[..., off:off + w].reshape(head + shape)
        off += w
    return out


def _all_gather_packed(buf, world, group):
    """One collective: every rank contributes an identically shaped buffer."""
    out = torch.empty((world,) + tuple(buf.shape), dtype=buf.dtype, device=buf.device)
    try:
        dist.all_gather_into_tensor(out.view(-1), buf.reshape(-1), group=group)
    except (RuntimeError, NotImplementedError):      # backends without the flat variant
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf.contiguous(), group=group)
        out = torch.stack(parts)
    return out


def gather_results(local, counts, group=None, num_grasps=None, n_objects=None):
    """local: dict of per-object tensors (dim 0 = this rank's objects; with fewer objects than ranks dim 1 = this rank's
    grasps).  Returns the dict over all objects / grasps, on every rank.  All tensors of one dtype travel in ONE padded
    buffer and ONE all_gather (the result dictionary of generate_grasps is all fp32: a single collective per batch).
    `counts` = objects per rank; with fewer objects than ranks pass num_grasps and n_objects as well and the grasp slices
    are put back together per object."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    split = n_objects is not None and 0 < n_objects < world
    lead = 2 if split else 1
    plans = [plan(n_objects, num_grasps, world, r) for r in range(world)] if split else None
    out = {}
    for dtype in sorted({v.dtype for v in local.values()}, key=str):
        keys = sorted(k for k, v in local.items() if v.dtype == dtype)
        shapes = [tuple(local[k].shape[lead:]) for k in keys]
        buf = _pack([local[k] for k in keys], lead)
        if split:
            g_max = max(p[3] - p[2] for p in plans)
            pad = torch.zeros((1, g_max, buf.shape[2]), dtype=dtype, device=buf.device)
            pad[:, :buf.shape[1]] = buf
            allb = _all_gather_packed(pad, world, group)                              # [world, 1, g_max, F]
            full = torch.cat([torch.cat([allb[r, :, :p[3] - p[2]] for r, p in enumerate(plans) if p[0] == o], dim=1)
                              for o in range(n_objects)], dim=0)
        elif len(set(counts)) == 1:
            # even shards (the usual case): no padding, and the gathered buffer already is the result
            full = _all_gather_packed(buf.contiguous(), world, group).view((world * counts[0],) + tuple(buf.shape[1:]))
        else:
            pad = torch.zeros((max(counts),) + tuple(buf.shape[1:]), dtype=dtype, device=buf.device)
            pad[:buf.shape[0]] = buf
            allb = _all_gather_packed(pad, world, group)                              # [world, max count, F]
            full = torch.cat([allb[r, :c] for r, c in enumerate(counts)], dim=0)
        out.update(_unpack(full, keys, shapes, lead))
    return out


def generate_sharded(generate_fn, pcs, num_grasps, rank=None, world_size=None, gather=True, group=None):
    """generate_fn(local_pcs, first_object_index[, num_grasps_local, first_grasp_index]) -> dict of per-object tensors
    ([n_local, G_local, ...]).  pcs: [n_objects, N, 3] (every rank holds, or can produce, the full list; only its slice
    is touched).  The two extra arguments are passed only when the grasps of an object are split (objects < ranks)."""
    rank = dist.get_rank(group) if rank is None else rank
    world_size = dist.get_world_size(group) if world_size is None else world_size
    n_obj = pcs.shape[0]
    lo, hi, g_lo, g_hi = plan(n_obj, num_grasps, world_size, rank)
    split = 0 < n_obj < world_size
    local = generate_fn(pcs[lo:hi], lo, g_hi - g_lo, g_lo) if split else generate_fn(pcs[lo:hi], lo)
    if not gather or world_size == 1:
        return local
    return gather_results(local, shard_counts(n_obj, world_size), group, num_grasps=num_grasps, n_objects=n_obj)
