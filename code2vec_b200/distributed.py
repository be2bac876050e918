"""Data-parallel plumbing for the training loop (BASELINE.json north_star: "training shards the corpus
across the 8 GPUs of one box with a single NCCL allreduce on the gradients per step over NVLink").

The reference has no distributed code at all (single device, main.py:83); this adds
the minimum: bags are independent, so the forward shards with no collective; only the gradients meet.

  shard_range / shard_items : contiguous shard of the methods of the corpus for a rank
  FlatGradBucket            : every parameter's .grad is a view into ONE flat fp32 buffer, so the step
                              needs exactly one all_reduce (NCCL over NVLink/NVSwitch; gloo in the CPU tests)
  ddp_step                  : zero_grad -> forward -> loss -> backward -> allreduce -> optimizer.step,
                              the body of main.py:171-175 with the collective inserted
torch.distributed is only plumbing here (process group, the collective); launch one process per GPU with
`python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 ...`.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """[lo, hi) of a contiguous, balanced split of n_items (first n_items % world ranks get one more)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_items(items, rank=None, world=None):
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_range(len(items), rank, world)
    return items[lo:hi]


class FlatGradBucket:
    """All gradients of `params` live in one flat buffer; `allreduce()` is the step's single collective."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=dt, device=dev)
        o = 0
        for p in self.params:
            if p.device != dev or p.dtype != dt:
                raise ValueError("parameters must share device and dtype")
            p.grad = self.flat[o:o + p.numel()].view_as(p)     # a view: backward accumulates in place
            o += p.numel()

    def zero(self):
        """optimizer.zero_grad(set_to_none=False) for the whole model in one memset (main.py:171)."""
        self.flat.zero_()

    def check_views(self):
        """.grad must still alias the flat buffer (zero_grad(set_to_none=True) would break that)."""
        o = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + o * self.flat.element_size():
                raise RuntimeError("a .grad no longer aliases the flat bucket; use bucket.zero(), not "
                                   "optimizer.zero_grad(set_to_none=True)")
            o += p.numel()

    def nbytes(self):
        return self.numel * self.flat.element_size()

    def allreduce(self, average=True):
        """ONE all_reduce(sum) over the flat bucket, then 1/world (mean of per-rank mean losses ==
        mean over the global batch when every rank holds the same number of bags)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            if average:
                self.flat.mul_(1.0 / dist.get_world_size())


def ddp_step(model, optimizer, bucket, starts, paths, ends, label, loss_fn):
    """One training step of main.py:171-175 on this rank's shard of the global batch."""
    bucket.zero()
    outputs, code_vector, attention = model.forward(starts, paths, ends, label)
    loss = loss_fn(outputs, label)
    loss.backward()
    bucket.allreduce()
    optimizer.step()
    return loss


def broadcast_parameters(model, src=0):
    """Same initial weights everywhere (the reference seeds one process; here rank `src` wins)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src)
