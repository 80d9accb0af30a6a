"""Learning-rate schedules of the path with the reference's checkpoint layout.

The reference builds a stock torch scheduler (``CosineAnnealingLR`` / ``MultiStepLR``,
fullbatch/training/optimizers.py:69-87) and, for ``hyp.warmup > 0``, wraps it in the third-party
``GradualWarmupScheduler`` (optimizers.py:89-91; fullbatch/training/additional_optimizers/scheduler.py:32-111, MIT,
ildoonet/pytorch-gradual-warmup-lr).  Checkpoints store ``scheduler.state_dict()`` (training/utils.py:43-51), so a
drop-in must produce and accept the same dictionaries: the torch classes are used as they are, and ``LinearWarmup`` below
is a fresh implementation of the wrapper's behaviour for ``multiplier == 1`` with the SAME attribute names (the state
dict is the instance ``__dict__`` minus the optimizer, with the wrapped scheduler's ``__dict__`` under
``'after_scheduler'``).

Schedule with warm-up W (scheduler.py:49-66): lr(t) = base * t / W for t <= W (so lr(0) = 0: the regulariser's lr/4
factor vanishes at step 0, modules.py:214), then the wrapped schedule starts at its epoch 0 at t = W + 1.
"""
import torch
from torch.optim.lr_scheduler import LRScheduler


class LinearWarmup(LRScheduler):
    def __init__(self, optimizer, total_epoch, after_scheduler, multiplier=1.0):
        if multiplier != 1.0:
            raise ValueError("only multiplier == 1.0 (lr rises from 0 to the base lr) is on the B200 path")
        self.multiplier = multiplier
        self.total_epoch = total_epoch
        self.after_scheduler = after_scheduler
        self.finished = False
        super().__init__(optimizer)

    def get_lr(self):
        if self.last_epoch > self.total_epoch:
            if not self.finished:  # hand over: the wrapped schedule starts from the base learning rates
                self.after_scheduler.base_lrs = [lr * self.multiplier for lr in self.base_lrs]
                self.finished = True
            return self.after_scheduler.get_last_lr()
        return [lr * float(self.last_epoch) / self.total_epoch for lr in self.base_lrs]

    def step(self, epoch=None):
        if self.finished:
            self.after_scheduler.step() if epoch is None else self.after_scheduler.step(epoch - self.total_epoch)
            self._last_lr = self.after_scheduler.get_last_lr()
            return None
        return super().step() if epoch is None else super().step(epoch)

    def state_dict(self):
        state = {k: v for k, v in self.__dict__.items() if k != "optimizer"}
        state["after_scheduler"] = {k: v for k, v in self.after_scheduler.__dict__.items() if k != "optimizer"}
        return state

    def load_state_dict(self, state_dict):
        state = dict(state_dict)
        self.after_scheduler.__dict__.update(state.pop("after_scheduler"))
        self.__dict__.update(state)


def build_scheduler(optimizer, cfg_hyp):
    """optimizers.py:69-91 for the schedules the path's configs use (cosine-4000 / cosine-decay / cosine-decay-floored /
    linear / exponential / none)."""
    sched = cfg_hyp.scheduler
    steps = int(cfg_hyp.steps)
    if sched == "linear":
        after = torch.optim.lr_scheduler.MultiStepLR(
            optimizer, milestones=[steps // 2.667, steps // 1.6, steps // 1.142], gamma=0.1)  # floats, as in the reference
    elif sched == "exponential":
        after = torch.optim.lr_scheduler.ExponentialLR(optimizer, gamma=0.99)
    elif sched == "cosine-decay-floored":
        after = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, steps, eta_min=cfg_hyp.optim.lr / 25)
    elif sched == "cosine-decay":
        after = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, steps, eta_min=0.0)
    elif sched == "cosine-4000":
        after = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, 4000, eta_min=0.0)
    elif sched in ["", " ", None]:
        after = torch.optim.lr_scheduler.MultiStepLR(optimizer, milestones=[], gamma=1)
    else:
        raise ValueError(f"Invalid scheduler {sched} provided.")
    warmup = int(cfg_hyp.warmup or 0)
    if warmup > 0:
        return LinearWarmup(optimizer, total_epoch=warmup, after_scheduler=after)
    return after
