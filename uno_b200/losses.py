"""Relative Lp loss used by the reference training loops (utilities3.py:75-103, LpLoss.rel)."""
import torch


class LpLoss:
    def __init__(self, d=2, p=2, size_average=True, reduction=True):
        assert d > 0 and p > 0
        self.d, self.p, self.reduction, self.size_average = d, p, reduction, size_average

    def rel(self, x, y):
        n = x.size()[0]
        diff = torch.norm(x.reshape(n, -1) - y.reshape(n, -1), self.p, 1)
        ynorm = torch.norm(y.reshape(n, -1), self.p, 1)
        if self.reduction:
            return torch.mean(diff / ynorm) if self.size_average else torch.sum(diff / ynorm)
        return diff / ynorm

    def __call__(self, x, y):
        return self.rel(x, y)
