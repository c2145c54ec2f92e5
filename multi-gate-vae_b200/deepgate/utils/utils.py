"""zero_normalization and AverageMeter (reference utils/utils.py:14-36)."""
import torch


class AverageMeter(object):
    """Running mean of a scalar."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        if self.count > 0:
            self.avg = self.sum / self.count


def zero_normalization(x):
    """(x - mean) / unbiased std.  The training loss uses the fused kernel (ops.vae_func_loss);
    this stays for callers that normalise label tensors themselves."""
    return (x - torch.mean(x)) / torch.std(x)
