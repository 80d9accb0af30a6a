"""ORACLE (test infrastructure only) -- restatement of the reference full-batch grad-reg step.

This file is a CHECKER.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under ``fullbatchtraining_b200/``
imports it, and the product path has no CPU fallback.

It restates, with plain ``torch`` tensor ops (``F.conv2d`` / ``F.batch_norm`` / autograd), the algorithm of
JonasGeiping/fullbatchtraining's hot path.  Every function cites the reference file:line it follows
(paths relative to the reference tree):

* model forward               fullbatch/models/resnets.py:43-126,128-177,179-230,271-316
* model factory quirks        fullbatch/models/models.py:14-22  (zero_init_residual is False on the hydra path)
* loss                        fullbatch/models/modules.py:96-101
* per-microbatch gradient     fullbatch/training/training.py:76-83
* forward-difference penalty  fullbatch/models/modules.py:211-241
* running-mean accumulation   fullbatch/training/training.py:45-47,121-185

Parity status: PINNED.  ``oracle/make_goldens.py`` runs the real reference (imported from /root/reference with the
shims in ``oracle/reference_harness.py``) in fp32 and fp64 and commits fingerprints under ``tests/golden``;
``tests/test_oracle_golden.py`` checks this restatement against them.

The third-party arithmetic itself (conv/BN/autograd) is PyTorch's (reference requires torch>=1.9; here
torch 2.11.0, oneDNN on CPU); fp64 runs of this oracle are the ground truth the CUDA path is judged against,
fp32 runs give the reference's own noise floor.
"""
from collections import OrderedDict
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # torch.nn.BatchNorm2d default, used by resnets.py:71 via get_layer_functions
BN_MOMENTUM = 0.1


def resnet_layout(depth):
    """(block kind, blocks per stage) -- resnets.py:12-40."""
    table = {
        18: ("basic", [2, 2, 2, 2]),
        34: ("basic", [3, 4, 6, 3]),
        50: ("bottleneck", [3, 4, 6, 3]),
        101: ("bottleneck", [3, 4, 23, 3]),
        152: ("bottleneck", [3, 8, 36, 3]),
    }
    return table[depth]


def build_resnet_state(depth, channels=3, classes=10, dtype=torch.float32):
    """Create parameters + buffers of the reference ResNet (CIFAR stem, downsample 'C') with the same RNG consumption
    order as ``ResNet.__init__`` (resnets.py:45-126), so that ``torch.manual_seed(s)`` followed by this call gives the
    same values as ``torch.manual_seed(s); construct_model(...)`` of the reference.

    Returns (params, buffers): OrderedDicts keyed like the reference ``state_dict``; ``params`` is in
    ``model.parameters()`` order (= the flat-buffer order of training/utils.py:34).
    """
    kind, layers = resnet_layout(depth)
    expansion = 1 if kind == "basic" else 4
    mods = OrderedDict()  # registration order == modules() order == parameters() order

    def conv(cin, cout, k, stride=1, padding=0):
        return torch.nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=padding, bias=False)

    inplanes = 64
    # stem: resnets.py:68-73
    mods["stem.0"] = conv(channels, inplanes, 3, 1, 1)
    mods["stem.1"] = torch.nn.BatchNorm2d(inplanes)
    width = inplanes
    strides = [1, 2, 2, 2]
    for s, nblocks in enumerate(layers):
        planes = width
        stride = strides[s]
        # _make_layer constructs the downsample op BEFORE the block (resnets.py:136-171)
        ds = None
        if stride != 1 or inplanes != planes * expansion:
            ds = (conv(inplanes, planes * expansion, 1), torch.nn.BatchNorm2d(planes * expansion))
        for b in range(nblocks):
            pre = f"layers.{s}.{b}"
            st = stride if b == 0 else 1
            if kind == "basic":
                mods[pre + ".conv1"] = conv(inplanes, planes, 3, st, 1)
                mods[pre + ".bn1"] = torch.nn.BatchNorm2d(planes)
                mods[pre + ".conv2"] = conv(planes, planes, 3, 1, 1)
                mods[pre + ".bn2"] = torch.nn.BatchNorm2d(planes)
            else:
                mods[pre + ".conv1"] = conv(inplanes, planes, 1)
                mods[pre + ".bn1"] = torch.nn.BatchNorm2d(planes)
                mods[pre + ".conv2"] = conv(planes, planes, 3, st, 1)
                mods[pre + ".bn2"] = torch.nn.BatchNorm2d(planes)
                mods[pre + ".conv3"] = conv(planes, planes * expansion, 1)
                mods[pre + ".bn3"] = torch.nn.BatchNorm2d(planes * expansion)
            if b == 0 and ds is not None:
                mods[pre + ".downsample.1"] = ds[0]
                mods[pre + ".downsample.2"] = ds[1]
            inplanes = planes * expansion
        width *= 2
    mods["fc"] = torch.nn.Linear(inplanes, classes)

    # init loop: resnets.py:109-114 (modules() order); zero_init_residual False (models.py:22 quirk)
    for m in mods.values():
        if isinstance(m, torch.nn.Conv2d):
            torch.nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        elif isinstance(m, torch.nn.BatchNorm2d):
            torch.nn.init.constant_(m.weight, 1)
            torch.nn.init.constant_(m.bias, 0)

    params, buffers = OrderedDict(), OrderedDict()
    for name, m in mods.items():
        for pn, p in m.named_parameters():
            params[f"{name}.{pn}"] = p.detach().to(dtype).clone()
        for bn, b in m.named_buffers():
            buffers[f"{name}.{bn}"] = b.detach().clone() if b.dtype == torch.long else b.detach().to(dtype).clone()
    return params, buffers


class OracleResNet:
    """Functional train-mode forward of the reference ResNet over a dict of parameters.

    resnets.py:179-189 (_forward_impl), :214-230 (BasicBlock.forward), :296-316 (Bottleneck.forward),
    :147-152 (downsample 'C' = AvgPool2d(stride) -> conv1x1 -> BN).
    """

    def __init__(self, depth, buffers, update_running_stats=True):
        self.kind, self.layers = resnet_layout(depth)
        self.buffers = buffers
        self.update_running_stats = update_running_stats

    def _bn(self, x, p, name):
        rm = self.buffers[name + ".running_mean"] if self.update_running_stats else None
        rv = self.buffers[name + ".running_var"] if self.update_running_stats else None
        if self.update_running_stats:
            self.buffers[name + ".num_batches_tracked"] += 1
        if rm is not None and rm.dtype != x.dtype:
            rm = rv = None  # fingerprints of running stats are only taken in matching dtype runs
        return F.batch_norm(x, rm, rv, p[name + ".weight"], p[name + ".bias"], True, BN_MOMENTUM, BN_EPS)

    def forward(self, p, x):
        out = F.conv2d(x, p["stem.0.weight"], None, 1, 1)
        out = F.relu(self._bn(out, p, "stem.1"))
        strides = [1, 2, 2, 2]
        for s, nblocks in enumerate(self.layers):
            for b in range(nblocks):
                pre = f"layers.{s}.{b}"
                st = strides[s] if b == 0 else 1
                identity = out
                if self.kind == "basic":
                    y = F.conv2d(out, p[pre + ".conv1.weight"], None, st, 1)
                    y = F.relu(self._bn(y, p, pre + ".bn1"))
                    y = F.conv2d(y, p[pre + ".conv2.weight"], None, 1, 1)
                    y = self._bn(y, p, pre + ".bn2")
                else:
                    y = F.conv2d(out, p[pre + ".conv1.weight"], None, 1, 0)
                    y = F.relu(self._bn(y, p, pre + ".bn1"))
                    y = F.conv2d(y, p[pre + ".conv2.weight"], None, st, 1)
                    y = F.relu(self._bn(y, p, pre + ".bn2"))
                    y = F.conv2d(y, p[pre + ".conv3.weight"], None, 1, 0)
                    y = self._bn(y, p, pre + ".bn3")
                if (pre + ".downsample.1.weight") in p:
                    identity = F.avg_pool2d(out, st, st) if st > 1 else out  # AvgPool2d(1,1) is the identity
                    identity = F.conv2d(identity, p[pre + ".downsample.1.weight"], None, 1, 0)
                    identity = self._bn(identity, p, pre + ".downsample.2")
                out = F.relu(y + identity)
        out = out.mean(dim=(2, 3))  # AdaptiveAvgPool2d((1,1)) + flatten
        return F.linear(out, p["fc.weight"], p["fc.bias"])


def label_smooth_xent(logits, target, smoothing=0.0):
    """modules.py:96-101."""
    log_prob = F.log_softmax(logits, dim=-1)
    weight = torch.ones_like(logits) * smoothing / (logits.shape[-1] - 1.0)
    weight.scatter_(-1, target.unsqueeze(-1), (1.0 - smoothing))
    return (-weight * log_prob).sum(dim=-1).mean()


def microbatch_gradient(net, p, x, y, smoothing=0.0):
    """training.py:76-83: loss, correct count, gradient list (parameters() order)."""
    names = list(p.keys())
    leaves = OrderedDict((k, v.detach().requires_grad_(True)) for k, v in p.items())
    with torch.enable_grad():
        logits = net.forward(leaves, x)
        loss = label_smooth_xent(logits, y, smoothing)
    correct = (logits.argmax(dim=-1) == y).float().sum()
    grads = torch.autograd.grad(loss, [leaves[k] for k in names])
    return [g.detach() for g in grads], loss.detach(), correct.detach()


@torch.no_grad()
def regularize(net, p, grads, x, y, lr, block_strength, eps, smoothing=0.0, implementation="forward-differences",
               pre_grads=None, acc_strength=0.0):
    """GradRegularizer.forward (modules.py:151-175 dispatch): returns (regularised grads, eps_n, hvp); ``p`` unchanged.

    forward-differences         modules.py:211-241
    forward-differences-legacy  modules.py:243-264  (v = g, cf *= block_strength: the same update up to rounding)
    central-differences         modules.py:266-300  (theta +- 0.5*eps_n*v, three gradient evaluations in total)
    acc_strength != 0 adds acc_strength * pre_grads to the direction v (modules.py:220-221, :273-274).
    """
    cf = lr / 4  # :214
    names = list(p.keys())
    if implementation == "forward-differences-legacy":
        vec = [g.clone() for g in grads]  # :247 (pre_grads are disregarded, :244)
        cf = cf * block_strength  # :246
    else:
        vec = [g * block_strength for g in grads]  # :217
        if pre_grads is not None:
            vec = [v + acc_strength * q for v, q in zip(vec, pre_grads)]  # :220-221
    eps_n = eps / torch.stack([v.pow(2).sum() for v in vec]).sum().sqrt()  # :223
    if implementation in ("forward-differences", "forward-differences-legacy", "finite_diff"):
        shifted = OrderedDict((k, p[k] + eps_n * v) for k, v in zip(names, vec))  # :226
        g2, _, _ = microbatch_gradient(net, shifted, x, y, smoothing)  # :227-230
        hvp = [(b - a) / eps_n for a, b in zip(grads, g2)]  # :232-234
    elif implementation == "central-differences":
        plus = OrderedDict((k, p[k] + 0.5 * eps_n * v) for k, v in zip(names, vec))  # :279
        gp, _, _ = microbatch_gradient(net, plus, x, y, smoothing)
        minus = OrderedDict((k, plus[k] - eps_n * v) for k, v in zip(names, vec))  # :286 (applied to the shifted params)
        gm, _, _ = microbatch_gradient(net, minus, x, y, smoothing)
        hvp = [(a - b) / eps_n for a, b in zip(gp, gm)]  # :292-293
    else:
        raise ValueError(f"Invalid spec. given for regularizer implementation: {implementation}")  # :175
    out = [a + cf * h for a, h in zip(grads, hvp)]  # :240 / :299
    return out, eps_n, hvp


def forward_differences(net, p, grads, x, y, lr, block_strength, eps, smoothing=0.0):
    """modules.py:211-241 with acc_strength == 0."""
    return regularize(net, p, grads, x, y, lr, block_strength, eps, smoothing, "forward-differences")


def clip_gradient_list(grads, clip, eps=1e-6):
    """training/utils.py:4-19 with grad_clip_norm = 2: in-place-equivalent global L2 clip; returns (grads, clipped?)."""
    norm = torch.norm(torch.stack([torch.norm(g, 2) for g in grads]), 2)
    if norm > clip:
        return [g * (clip / (norm + eps)) for g in grads], 1
    return grads, 0


@torch.no_grad()
def full_batch_step(depth, p, buffers, X, Y, mb, lr, block_strength=0.5, eps=1e-2, smoothing=0.0,
                    acc_dtype=None, keep_microbatches=0, order=None, implementation="forward-differences",
                    acc_strength=0.0, batch_clip=None):
    """training.py:121-185, single process (num_machines = 1).

    X: [N,3,32,32], Y: [N] int64.  Microbatches are consecutive blocks of ``mb`` images, drop_last
    (data_preparation.py:56-72), optionally permuted by ``order`` (a permutation of the sample indices, hyp.shuffle).
    Returns dict with avg gradient list, mean loss, correct count, grad_norms (squared, raw) and optionally the
    first ``keep_microbatches`` raw / regularised microbatch gradients.
    """
    net = OracleResNet(depth, buffers)
    acc_dtype = acc_dtype or X.dtype
    K = X.shape[0] // mb

    def batch(k):
        idx = slice(k * mb, (k + 1) * mb) if order is None else order[k * mb:(k + 1) * mb]
        return X[idx], Y[idx]

    pre_grads = None
    if acc_strength != 0:  # training.py:128-142: full extra sweep for the mean raw gradient
        pre_grads = [torch.zeros_like(v, dtype=acc_dtype) for v in p.values()]
        for k in range(K):
            x, y = batch(k)
            g, _, _ = microbatch_gradient(net, p, x, y, smoothing)
            g = [t.to(acc_dtype) for t in g]
            if batch_clip is not None:
                g, _ = clip_gradient_list(g, batch_clip)
            for a, t in zip(pre_grads, g):
                a.add_(t - a, alpha=1 / (k + 1))
    avg = [torch.zeros_like(v, dtype=acc_dtype) for v in p.values()]  # :123
    grad_norms = torch.zeros(K, dtype=X.dtype, device=X.device)
    step_loss = torch.zeros((), dtype=X.dtype, device=X.device)
    step_preds = torch.zeros((), dtype=X.dtype, device=X.device)
    kept = []
    clipped_batches = 0
    for k in range(K):
        x, y = batch(k)
        g, loss, correct = microbatch_gradient(net, p, x, y, smoothing)  # :159
        grad_norms[k] = torch.stack([t.pow(2).sum() for t in g]).sum()  # :162
        if block_strength != 0 or acc_strength != 0:
            g_reg, eps_n, _ = regularize(net, p, g, x, y, lr, block_strength, eps, smoothing, implementation,
                                         pre_grads, acc_strength)  # :163
        else:
            g_reg, eps_n = g, None  # modules.py:151-153,177-178 (_pass)
        if k < keep_microbatches:
            kept.append(dict(raw=g, reg=g_reg, loss=loss, correct=correct, eps_n=eps_n))
        g_acc = [t.to(acc_dtype) for t in g_reg]  # :165
        if batch_clip is not None:  # :166-167
            g_acc, c = clip_gradient_list(g_acc, batch_clip)
            clipped_batches += c
        for a, t in zip(avg, g_acc):  # :45-47,168  (avg += (g - avg) / (k+1))
            t = t - a
            a.add_(t, alpha=1 / (k + 1))
        step_loss += loss
        step_preds += correct
    param_norm = sum(v.pow(2).sum() for v in p.values())
    return dict(avg=avg, loss=step_loss / K, correct=step_preds, grad_norms=grad_norms, kept=kept, K=K,
                param_norm=param_norm, pre_grads=pre_grads, clipped_batches=clipped_batches)


def flat(tensors):
    return torch.cat([t.reshape(-1) for t in tensors])


def fingerprint(tensors, stride=997):
    """Compact, machine-independent fingerprint of a list of tensors: per-tensor L2 norms and sums, plus a strided
    sample of the flat vector.  Used for the committed goldens (full vectors are 45 MB each)."""
    f = flat(tensors).double().cpu()
    return dict(
        norms=torch.stack([t.double().norm() for t in tensors]).cpu().numpy(),
        sums=torch.stack([t.double().sum() for t in tensors]).cpu().numpy(),
        sample=f[::stride].numpy().copy(),
        total_norm=float(f.norm()),
    )


def synthetic_cifar(n, seed=1234, dtype=torch.float32):
    """randn images (~normalised CIFAR) and uniform labels from one seeded generator."""
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, 32, 32, generator=gen)
    y = torch.randint(0, 10, (n,), generator=gen)
    return x.to(dtype), y


@torch.no_grad()
def sgd_epochs(depth, p, buffers, X, Y, mb, steps, lr, momentum=0.9, weight_decay=5e-4, nesterov=True,
               block_strength=0.0, eps=1e-2, smoothing=0.0, grad_clip=None, implementation="forward-differences"):
    """The stochastic sanity branch, training.py:241-286, single process: per step one pass over the loader blocks
    (sequential order, drop_last), per block raw gradient (:258) -> squared norm (:259) -> GradRegularizer (:260) ->
    optional clip_grad_norm_ (:271-272) -> torch.optim.SGD step (optimizers.py:25-28; torch/optim/sgd.py: d = g + wd*theta;
    buf = d on first use else momentum*buf + d; d = d + momentum*buf (Nesterov); theta -= lr*d); the cosine-4000
    scheduler steps once per step (:284; no warm-up).  `p` is updated IN PLACE.  Returns per-step train_loss / train_acc
    / mean squared raw gradient norm."""
    import math

    net = OracleResNet(depth, buffers)
    K = X.shape[0] // mb
    bufs = None
    out = dict(train_loss=[], train_acc=[], grad_norm_sq=[])
    for step in range(steps):
        cur_lr = lr * (1 + math.cos(math.pi * step / 4000)) / 2
        loss_sum, preds, norms = 0.0, 0.0, []
        for k in range(K):
            x, y = X[k * mb:(k + 1) * mb], Y[k * mb:(k + 1) * mb]
            g, loss, correct = microbatch_gradient(net, p, x, y, smoothing)
            norms.append(float(torch.stack([t.pow(2).sum() for t in g]).sum()))
            if block_strength != 0:
                g, _, _ = regularize(net, p, g, x, y, cur_lr, block_strength, eps, smoothing, implementation, None, 0.0)
            if grad_clip is not None:  # torch.nn.utils.clip_grad_norm_: coef = clip / (norm + 1e-6), clamped to 1
                total = torch.norm(torch.stack([torch.norm(t, 2) for t in g]), 2)
                coef = torch.clamp(grad_clip / (total + 1e-6), max=1.0)
                g = [t * coef for t in g]
            d = [t + weight_decay * v for t, v in zip(g, p.values())]
            if momentum != 0:
                bufs = [t.clone() for t in d] if bufs is None else [b * momentum + t for b, t in zip(bufs, d)]
                d = [t + momentum * b for t, b in zip(d, bufs)] if nesterov else bufs
            for v, t in zip(p.values(), d):
                v.sub_(cur_lr * t)
            loss_sum += float(loss)
            preds += float(correct)
        out["train_loss"].append(loss_sum / K)
        out["train_acc"].append(preds / (K * mb))
        out["grad_norm_sq"].append(sum(norms) / K)
    return out


def structured_cifar(n, seed=4321, dtype=torch.float32):
    """Image-like synthetic data: per class a smooth oriented pattern, plus a low-frequency random field (bilinear
    upsampling of 4x4 noise) and a little pixel noise, standardised per channel -- spatially correlated inputs with
    class structure, unlike the white noise of synthetic_cifar."""
    import math

    gen = torch.Generator().manual_seed(seed)
    y = torch.randint(0, 10, (n,), generator=gen)
    coarse = torch.randn(n, 3, 4, 4, generator=gen)
    field = torch.nn.functional.interpolate(coarse, size=(32, 32), mode="bilinear", align_corners=False)
    noise = 0.1 * torch.randn(n, 3, 32, 32, generator=gen)
    yy, xx = torch.meshgrid(torch.arange(32.0), torch.arange(32.0), indexing="ij")
    angle = (y.float() * math.pi / 10)[:, None, None]
    freq = (1 + (y % 3).float())[:, None, None] * 2 * math.pi / 32
    wave = torch.sin(freq * (xx[None] * torch.cos(angle) + yy[None] * torch.sin(angle)))
    phase = torch.tensor([0.0, 0.7, 1.9])[None, :, None, None]
    x = field + noise + 0.8 * torch.sin(torch.asin(wave.clamp(-1, 1))[:, None] + phase)
    x = (x - x.mean(dim=(0, 2, 3), keepdim=True)) / x.std(dim=(0, 2, 3), keepdim=True)
    return x.to(dtype), y


def flops_per_image(depth):
    """Algorithmic GFLOP per image for one grad-reg step (2 passes); BASELINE.md section 2."""
    return {18: 6.6580, 152: 44.6586}[depth]
