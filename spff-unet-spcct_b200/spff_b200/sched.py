"""Learning-rate schedule of the hot path's optimizer, host side.

The reference steps `torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode='max', factor=0.5, patience=5)`
on `val_macro_dice` once per validation epoch (innovative3D/models.py:591-594). The fused Adam kernel
takes the learning rate as a plain argument, so the schedule is a few lines of host arithmetic with the
same rules (relative threshold 1e-4, no cooldown, eps 1e-8)."""
from __future__ import annotations


class PlateauLR:
    def __init__(self, lr: float, mode: str = "max", factor: float = 0.5, patience: int = 5, threshold: float = 1e-4,
                 min_lr: float = 0.0, eps: float = 1e-8):
        if mode not in ("max", "min"):
            raise ValueError("mode must be 'max' or 'min'")
        self.lr, self.mode, self.factor, self.patience = float(lr), mode, float(factor), int(patience)
        self.threshold, self.min_lr, self.eps = float(threshold), float(min_lr), float(eps)
        self.best = float("-inf") if mode == "max" else float("inf")
        self.num_bad = 0

    def _better(self, v: float) -> bool:
        if self.mode == "max":
            return v > self.best * (1.0 + self.threshold)
        return v < self.best * (1.0 - self.threshold)

    def step(self, metric: float) -> float:
        """Feed one validation metric; returns the (possibly reduced) learning rate."""
        v = float(metric)
        if self._better(v):
            self.best, self.num_bad = v, 0
        else:
            self.num_bad += 1
        if self.num_bad > self.patience:
            new = max(self.lr * self.factor, self.min_lr)
            if self.lr - new > self.eps:
                self.lr = new
            self.num_bad = 0
        return self.lr
