"""Data-parallel plumbing for the generator forward (one process per GPU, torch.distributed).

The path shards over independent images (no op mixes samples), so the only collective is one
all-gather of the packed output tiles ``[B/n, 4, H, W]`` (3 image channels + the soft mask). Backend-agnostic: NCCL
over NVLink on the GPU box, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_bounds(total, world, rank):
    """Contiguous, equal shards (total must divide evenly, like the reference's batchSize % n_gpu assert,
    reference options/base_options.py:180-183)."""
    if total % world:
        raise ValueError("global batch %d is not a multiple of world size %d" % (total, world))
    per = total // world
    return rank * per, (rank + 1) * per


def pack_outputs(composed, mask):
    return torch.cat([composed, mask], 1).contiguous()


def unpack_outputs(packed):
    return packed[:, :3], packed[:, 3:4]


def all_gather_outputs(composed, mask, out=None):
    """Single collective of the path: returns (composed_all, mask_all) ordered by rank."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    packed = pack_outputs(composed, mask)
    if world == 1:
        return unpack_outputs(packed)
    if out is None:
        out = packed.new_empty((world * packed.shape[0],) + tuple(packed.shape[1:]))
    dist.all_gather_into_tensor(out, packed)
    return unpack_outputs(out)
