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
    def __init__(self, params, bucket_bytes=64 << 20, process_group=None, hooks=True):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        order = list(reversed(self.params))
        # every tensor starts at a multiple of 128 elements: 16-byte alignment of fp32 AND bf16 views of the same layout
        # (cuBLAS drops to much slower kernels for unaligned operands)
        self.offsets = {}
        total = 0
        for p in order:
            self.offsets[p] = total
            total += (p.numel() + 127) // 128 * 128
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
