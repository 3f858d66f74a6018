"""Data-parallel plumbing for the ACR training step (one process per GPU, torch.distributed / NCCL).

The reference wraps the model in DistributedDataParallel but calls `model.module.forward_mirror`, which bypasses
the reducer, so no gradient is ever all-reduced (train_acr.py:99,138).  The north star asks for a real
all-reduce, so `GradBuckets` implements one: every trainable parameter's .grad is a VIEW into one flat fp32 buffer
(no flatten / unflatten copies); the buffer is cut into buckets in reverse registration order (the order gradients
become ready in backward); a post-accumulate hook launches an asynchronous NCCL all-reduce (AVG) of a bucket as
soon as its last gradient has landed, so communication over NVLink/NVSwitch overlaps the rest of the backward pass.
CAM inference shards images round-robin with no collective (`shard_indices`).
"""
import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world_size):
    """Round-robin image sharding for CAM inference (the reference makes every rank redo all images,
    infer_cam.py:119-128).  No collective is needed."""
    return list(range(rank, n_items, world_size))


class GradBuckets:
    def __init__(self, params, bucket_bytes=64 << 20, process_group=None, hooks=True, sharded=None):
        """sharded: optional collection of parameters whose optimiser state may be SHARDED across ranks (the trunk's Linear
        weights: the training step only reads their bf16 copies).  They are laid out first, in one region padded to a multiple
        of world x 128 elements (`shard_end`); `reduce_sharded()` then reduce-scatters that region and all-reduces the rest."""
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        order = list(reversed(self.params))
        self.shard_end = 0
        if sharded:
            ids = {id(p) for p in sharded}
            order = [p for p in order if id(p) in ids] + [p for p in order if id(p) not in ids]
            nshard = sum(1 for p in order if id(p) in ids)
        # every tensor starts at a multiple of 128 elements: 16-byte alignment of fp32 AND bf16 views of the same layout
        # (cuBLAS drops to much slower kernels for unaligned operands)
        self.offsets = {}
        total = 0
        for i, p in enumerate(order):
            if sharded and i == nshard:
                unit = 128 * self.world
                total = (total + unit - 1) // unit * unit        # every rank's shard starts on a 128-element boundary
                self.shard_end = total
            self.offsets[p] = total
            total += (p.numel() + 127) // 128 * 128
        if sharded and nshard == len(order):
            unit = 128 * self.world
            total = (total + unit - 1) // unit * unit
            self.shard_end = total
        self.grad_shard = None
        dev = order[0].device
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.buckets = []            # (start, end, n_params)
        self._bucket_of = {}
        off, start, count = 0, 0, 0
        for p in order:
            n = p.numel()
            off = self.offsets[p]
            p.grad = self.flat[off:off + n].view_as(p)
            self._bucket_of[p] = len(self.buckets)
            off += (n + 127) // 128 * 128
            count += 1
            if (off - start) * 4 >= bucket_bytes:
                self.buckets.append([start, off, count])
                start, count = off, 0
        if count:
            self.buckets.append([start, off, count])
        self._pending = [b[2] for b in self.buckets]
        self._works = []
        self._used = set()
        # gloo (CPU tests) has no AVG: sum, then scale after the wait
        self._avg = self.world > 1 and dist.get_backend(process_group) == "nccl"
        if self.world > 1 and hooks:
            for p in order:
                p.register_post_accumulate_grad_hook(self._hook)

    def _hook(self, p):
        b = self._bucket_of[p]
        self._pending[b] -= 1
        self._used.add(p)
        if self._pending[b] == 0:
            s, e, _ = self.buckets[b]
            self._works.append(self._reduce(s, e))

    def _reduce(self, s, e):
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        return dist.all_reduce(self.flat[s:e], op=op, group=self.group, async_op=True)

    def zero(self):
        self.flat.zero_()

    def reduce_all(self):
        """One all-reduce of the whole flat buffer (CUDA-graph mode: the backward runs as a graph replay, so there is
        no per-bucket hook to overlap with; 361 MB over NVLink/NVSwitch is ~0.6 ms)."""
        if self.world > 1:
            self._reduce(0, self.flat.numel()).wait()
            if not self._avg:
                self.flat.div_(self.world)

    def reduce_chunks(self, nchunks=4):
        """All-reduce the flat buffer as `nchunks` back-to-back collectives and yield (start, end) as each one becomes
        consumable: NCCL runs them in order on its own stream and `work.wait()` only makes the CURRENT stream wait, so the
        consumer (the optimiser update of that slice) runs under the all-reduce of the following slices."""
        n = self.flat.numel()
        step = ((n + nchunks - 1) // nchunks + 127) // 128 * 128
        spans = [(s, min(n, s + step)) for s in range(0, n, step)]
        if self.world == 1:
            yield from spans
            return
        works = [self._reduce(s, e) for s, e in spans]
        for (s, e), w in zip(spans, works):
            w.wait()
            if not self._avg:
                self.flat[s:e].div_(self.world)
            yield s, e

    def shard_range(self):
        """[start, end) of this rank's slice of the sharded region."""
        n = self.shard_end // self.world
        return self.rank * n, (self.rank + 1) * n

    def reduce_sharded(self):
        """Gradient exchange for a sharded optimiser: reduce-scatter (mean) of the sharded region -- this rank's slice of the
        averaged gradients lands in `grad_shard` -- and an all-reduce (mean) of the small replicated rest.  Moves (W-1)/W of the
        buffer once instead of twice.  gloo (CPU tests) has neither reduce_scatter nor AVG: all-reduce + slice there."""
        s, e = self.shard_range()
        if self.grad_shard is None:
            self.grad_shard = torch.empty(e - s, device=self.flat.device, dtype=torch.float32)
        if self.world == 1:
            self.grad_shard.copy_(self.flat[s:e])
            return
        big, small = self.flat[:self.shard_end], self.flat[self.shard_end:]
        if self._avg:
            w1 = dist.reduce_scatter_tensor(self.grad_shard, big, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            w2 = dist.all_reduce(small, op=dist.ReduceOp.AVG, group=self.group, async_op=True) if small.numel() else None
            w1.wait()
            if w2 is not None:
                w2.wait()
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(self.world)
            self.grad_shard.copy_(self.flat[s:e])

    def all_gather_shards(self, full):
        """In-place all-gather: every rank contributes its slice [shard_range) of `full[:shard_end]` (any dtype)."""
        if self.world == 1:
            return
        s, e = self.shard_range()
        region = full[:self.shard_end]
        if self._avg:       # nccl: the in-place form (send buffer = this rank's slice of the receive buffer)
            dist.all_gather_into_tensor(region, region[s:e], group=self.group)
        else:
            parts = [torch.empty(e - s, dtype=full.dtype, device=full.device) for _ in range(self.world)]
            dist.all_gather(parts, region[s:e].clone(), group=self.group)
            for r, part in enumerate(parts):
                region[r * (e - s):(r + 1) * (e - s)].copy_(part)

    def finish(self):
        """Call after backward: reduce buckets whose hooks did not all fire (parameters without gradient this step,
        e.g. norm/head/bkg_token) and wait for every outstanding all-reduce."""
        if self.world > 1:
            for b, left in enumerate(self._pending):
                if left > 0:
                    s, e, _ = self.buckets[b]
                    self._works.append(self._reduce(s, e))
            for w in self._works:
                w.wait()
            if not self._avg:
                self.flat.div_(self.world)
        self._works = []
        self._pending = [b[2] for b in self.buckets]
