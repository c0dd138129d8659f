"""Data-parallel plumbing for the generator forward (one process per GPU, torch.distributed).

The path shards over independent images (no op mixes samples), so the only collective is one
all-gather of the packed output tiles ``[B/n, 4, H, W]`` (3 image channels + the soft mask) per step. Backend-agnostic:
NCCL over NVLink on the GPU box, gloo in the CPU tests.

``OutputGather`` removes everything around that collective from the critical path: the heads of the forward write
straight into this rank's slice of the gather buffer (``Engine.inference_packed(out=slot)``: no pack / concat pass), the
all-gather is IN PLACE (send buffer = the slice) and asynchronous (``async_op=True``: it runs on the backend's own
stream), and two buffers alternate so the collective of step i overlaps the compute of step i+1.
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


def unpack_outputs(packed):
    return packed[:, :3], packed[:, 3:4]


class OutputGather:
    """Ring of ``depth`` gather buffers [world*B, 4, H, W] (fp32) on ``device``.

        slot = g.next_slot()                 # this rank's [B,4,H,W] slice of the next buffer (producer writes here)
        eng.inference_packed(img, sk, out=slot)
        g.launch()                           # async in-place all-gather of that buffer
        ...                                  # next step computes into the other buffer meanwhile
        full = g.wait()                      # all outstanding collectives done (stream-ordered on NCCL) -> last full buffer
    """

    def __init__(self, B, H, W, device, depth=2, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.B, self.depth = B, depth
        self.bufs = [torch.empty(self.world * B, 4, H, W, device=device, dtype=torch.float32) for _ in range(depth)]
        self.works = [None] * depth
        self.i = -1
        # NCCL gathers in place (send buffer == this rank's slice of the receive buffer); gloo needs a separate send buffer
        self.inplace = self.world > 1 and dist.get_backend(group) == "nccl"

    def _mine(self, k):
        return self.bufs[k][self.rank * self.B:(self.rank + 1) * self.B]

    def _finish(self, k):
        if self.works[k] is not None:
            self.works[k].wait()      # NCCL: the CURRENT stream waits for the collective (no host block)
            self.works[k] = None

    def next_slot(self):
        self.i += 1
        k = self.i % self.depth
        self._finish(k)               # the collective that last read / wrote this buffer
        return self._mine(k)

    def launch(self):
        k = self.i % self.depth
        if self.world > 1:
            send = self._mine(k) if self.inplace else self._mine(k).clone()
            self.works[k] = dist.all_gather_into_tensor(self.bufs[k], send, group=self.group, async_op=True)
        return k

    def wait(self, k=None):
        for j in (range(self.depth) if k is None else (k,)):
            self._finish(j)
        return self.bufs[(self.i if k is None else k) % self.depth] if self.i >= 0 else None


def all_gather_outputs(composed, mask):
    """One-shot form (tests, small callers): (composed_all, mask_all) ordered by rank."""
    B, _, H, W = composed.shape
    g = OutputGather(B, H, W, composed.device, depth=1)
    slot = g.next_slot()
    slot[:, :3].copy_(composed)
    slot[:, 3:4].copy_(mask)
    g.launch()
    return unpack_outputs(g.wait())
