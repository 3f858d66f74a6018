"""The ACR training step (train_acr.py:127-174) on this repo's kernels.

PolyOptimizer mirrors tool/torchutils.py:10-31 INCLUDING its argument slip: the reference passes
weight_decay positionally into SGD's momentum slot, so the effective optimiser is SGD(momentum=wt_dec, weight_decay=0)
with lr_t = lr * (1 - t/max_step)^0.9.
"""
import torch

from .losses import acr_total_loss
from .parallel import GradBuckets


class PolyOptimizer(torch.optim.SGD):
    def __init__(self, params, lr, weight_decay, max_step, momentum=0.9):
        super().__init__(params, lr, weight_decay)          # sic: lands in `momentum` (tool/torchutils.py:13)
        self.global_step = 0
        self.max_step = max_step
        self.momentum = momentum                             # the poly exponent (tool/torchutils.py:24)
        self.__initial_lr = [group['lr'] for group in self.param_groups]

    def step(self, closure=None):
        if self.global_step < self.max_step:
            lr_mult = (1 - self.global_step / self.max_step) ** self.momentum
            for i in range(len(self.param_groups)):
                self.param_groups[i]['lr'] = self.__initial_lr[i] * lr_mult
        super().step(closure)
        self.global_step += 1


class Trainer:
    """One object = one rank.  `step(img, label)` takes HOST (pinned) or device tensors and returns the loss tensor
    (device, detached); gradients are averaged across ranks when torch.distributed is initialised."""

    def __init__(self, model, lr=0.01, wt_dec=5e-4, max_step=100000, alpha=100.0, bucket_bytes=64 << 20):
        self.model = model
        self.alpha = alpha
        model.train()
        model.set_capture_grad(False)
        # parameters the ACR path never reaches get no gradient in the reference either
        self.buckets = GradBuckets(list(model.parameters()), bucket_bytes)
        self.opt = PolyOptimizer(model.parameters(), lr=lr, weight_decay=wt_dec, max_step=max_step)
        self.dev = next(model.parameters()).device
        self._img = None
        self._label = None

    def _stage(self, img, label):
        if img.is_cuda:
            return img, label
        if self._img is None or self._img.shape != img.shape:
            self._img = torch.empty(img.shape, device=self.dev, dtype=img.dtype)
            self._label = torch.empty(label.shape, device=self.dev, dtype=label.dtype)
        self._img.copy_(img, non_blocking=True)
        self._label.copy_(label, non_blocking=True)
        return self._img, self._label

    def step(self, img, label):
        img, label = self._stage(img, label)
        img2 = img.flip(-1)                                   # transforms.RandomHorizontalFlip(p=1), train_acr.py:135
        cls_list, (attn1, attn2) = self.model.forward_mirror(img, img2)
        loss, parts = acr_total_loss(cls_list[0], cls_list[1], label, attn1, attn2, img.shape[2] // 16, self.alpha)
        self.buckets.zero()
        loss.backward()
        self.buckets.finish()
        self.opt.step()
        return loss.detach()
