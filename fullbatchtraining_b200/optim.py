"""Fused optimizer step for the flat parameter buffer.

``FlatSGD`` is a ``torch.optim.Optimizer`` with the constructor, ``param_groups`` and ``state_dict`` layout of
``torch.optim.SGD`` (the reference's ``Gradient Descent`` optimizer, fullbatch/training/optimizers.py:25-28), so the
reference's schedulers and checkpoint format keep working, but ``step(closure)`` runs the global-norm clip of
``_modify_gradient_params`` (fullbatch/training/training.py:198-211), the SGD update and the ``sum theta^2`` of
``_record_stats`` (training.py:92) as one device sweep (``fb_sgd_step``) with no host synchronisation.
"""
import torch

from . import ops

S_GNORM, S_PNORM = 7, 8


class FlatSGD(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False):
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")  # same check as torch SGD
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("FlatSGD supports a single parameter group (the reference path uses one)")
        self.engine = None
        self.grad_clip = None
        self._buf = None
        self._first = True

    def bind(self, engine, grad_clip=None):
        """Attach the engine whose flat ``theta`` / ``avg`` buffers hold the parameters and the gradient."""
        self.engine = engine
        self.grad_clip = grad_clip
        group = self.param_groups[0]
        if [id(p) for p in group["params"]] != [id(p) for p in engine.model.parameters()]:
            raise ValueError("FlatSGD must own exactly model.parameters() in order")
        if group["momentum"] != 0:
            self._buf = torch.zeros_like(engine.theta)
            for p, v in zip(group["params"], engine.grads_list(self._buf)):
                self.state[p]["momentum_buffer"] = v  # torch.optim.SGD state layout (views of the flat buffer)
        return self

    def load_state_dict(self, state_dict):
        """torch.optim.SGD checkpoints (training/utils.py:53-70): momentum buffers are copied INTO the flat buffer so its
        per-parameter views stay the optimizer state."""
        super().load_state_dict(state_dict)
        if self._buf is not None and self.engine is not None:
            group = self.param_groups[0]
            for p, v in zip(group["params"], self.engine.grads_list(self._buf)):
                loaded = self.state[p].get("momentum_buffer")
                if loaded is not None:
                    if loaded.data_ptr() != v.data_ptr():
                        v.copy_(loaded.to(v.device))
                    self._first = False
                self.state[p]["momentum_buffer"] = v

    @torch.no_grad()
    def step(self, closure=None, grad=None):
        """closure() must leave the gradient in engine.avg (Trainer does); `grad` overrides the flat gradient buffer."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        eng, g = self.engine, self.param_groups[0]
        flat_grad = eng.avg if grad is None else grad
        clip = float(self.grad_clip) if self.grad_clip is not None else 0.0
        if clip > 0:
            ops.flat_sqnorm(flat_grad, eng.numel, eng.sq_ws, eng.scal, S_GNORM)
        ops.sgd_step(eng.theta, flat_grad, self._buf, eng.numel, eng.scal, S_GNORM, clip, g["lr"], g["momentum"],
                     g["dampening"], g["weight_decay"], g["nesterov"], self._first, clip > 0, eng.sq_ws, S_PNORM)
        self._first = False
        return loss
