"""Whole-microbatch executor of the full-batch gradient-regularised step on one B200.

Replaces the device work of ``_accumulate_full_gradient`` (reference fullbatch/training/training.py:121-185) and
``GradRegularizer._forward_differences`` (fullbatch/models/modules.py:211-241):

  per microbatch k (a replay of ONE captured CUDA graph, no host synchronisation):
    im2col of the microbatch (stem)                                              [fb_stem_im2col]
    pass 1: forward / loss / backward at theta           -> g      (flat fp32)   [tcgen05 convs + layer kernels]
    n2 = |g|^2 -> grad_norms[k]; eps_n = eps / (bs*sqrt(n2)); theta' = theta + eps_n*bs*g   [fb_flat_sqnorm, fb_fd_perturb]
    pass 2: forward / loss / backward at theta'          -> g2
    g_reg = g + (lr/4) (g2 - g)/eps_n ; avg += (g_reg - avg)/(k+1)               [fb_fd_combine]

Persistent state (allocated once, kernels never allocate):
  theta, theta', g, g2, avg : flat fp32 buffers in model.parameters() order (= training/utils.py:34 order);
                              the model's parameters are re-pointed to views of ``theta`` so optimizers update it in place
  per conv: bf16 hi/lo GEMM operands of the weights, refreshed from theta / theta' at the start of each pass
  per layer: fp32 conv output, bf16 hi/lo activation planes, fp32 activation gradient, bf16 output gradient

theta itself is never modified by the regulariser, so the reference's "restore from a clone" (modules.py:237-238) is
exact by construction.
"""
import os

import torch

from . import ops
from .models import ResNet, ResidualBlock

S_N2, S_EPS, S_LOSS, S_CORRECT, S_LOSS2, S_CORRECT2, S_CF = 0, 1, 2, 3, 4, 5, 6
S_VSQ, S_CLIPPED, S_REGSQ = 9, 10, 11  # (7, 8 are used by optim.FlatSGD)

# hyp.grad_reg.implementation -> device recipe.  `forward-differences-legacy` (modules.py:243-264) perturbs along g
# instead of bs*g and scales the correction by bs: the same update up to rounding, so it shares the forward recipe.
IMPLEMENTATIONS = {"finite_diff": "forward", "forward-differences": "forward", "forward-differences-legacy": "forward",
                   "central-differences": "central"}
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


class Act:
    """An activation tensor [n,h,w,c]: bf16 hi/lo planes (+ fp32 gradient buffer when it needs one)."""

    def __init__(self, n, h, w, c, split, device, grad=True):
        self.n, self.h, self.w, self.c = n, h, w, c
        self.hi = torch.zeros(n, h, w, c, device=device, dtype=torch.bfloat16)
        self.lo = torch.zeros(n, h, w, c, device=device, dtype=torch.bfloat16) if split else None
        self.grad = torch.zeros(n, h, w, c, device=device, dtype=torch.float32) if grad else None
        # second addend of the gradient (shortcut branch of the consuming block); consumers read grad + grad2, so no
        # GEMM epilogue ever has to read-modify-write
        self.grad2 = None

    def add_grad2(self):
        self.grad2 = torch.zeros_like(self.grad)
        return self.grad2

    @property
    def P(self):
        return self.n * self.h * self.w


class Unit:
    """conv (bias-free) + BatchNorm: buffers, descriptors and parameter offsets."""

    def __init__(self, eng, conv_name, bn_name, x, cout, k, stride, dx_accumulate=False, needs_dx=True, stem=False,
                 dx_target=None):
        dev, split = eng.device, eng.split
        self.conv_name, self.bn_name, self.x, self.stem = conv_name, bn_name, x, stem
        self.cin, self.cout, self.k, self.stride = x.c, cout, k, stride
        n, h, w = x.n, x.h, x.w
        self.ho, self.wo = h // stride, w // stride
        self.P = n * self.ho * self.wo
        self.y = torch.zeros(n, self.ho, self.wo, cout, device=dev)
        self.dy = torch.zeros(n, self.ho, self.wo, cout, device=dev, dtype=torch.bfloat16)
        self.mean = torch.zeros(cout, device=dev)
        self.rstd = torch.zeros(cout, device=dev)
        taps = k * k
        self.taps = taps
        bf = dict(device=dev, dtype=torch.bfloat16)
        # two sets of bf16 GEMM operands: [0] from theta (refreshed once per step), [1] from theta' (every microbatch)
        self.w = []
        for _ in range(2):
            self.w.append((torch.zeros(cout, taps * x.c, **bf),
                           torch.zeros(cout, taps * x.c, **bf) if split else None,
                           torch.zeros(x.c, taps * cout, **bf) if needs_dx else None,
                           torch.zeros(x.c, taps * cout, **bf) if (needs_dx and split) else None))
        need = ops.Conv2dPlan.partial_elems(n, h, w, x.c, cout, k, stride)
        eng.partial_elems = max(eng.partial_elems, need)
        self.args = (n, h, w, x.c, cout, k, stride)
        self.dx_accumulate, self.needs_dx = dx_accumulate, needs_dx
        self.dx_target = dx_target  # fp32 buffer receiving the input gradient (default: x.grad)
        self.plans = None
        # unit whose BatchNorm(+ReLU) output is this unit's input and has no other consumer (conv1 -> conv2 inside a
        # block): this unit's dgrad epilogue then reduces that BatchNorm's backward statistics
        self.bn_producer = None

    def finish(self, eng):
        self.plans = [ops.Conv2dPlan(*self.args, self.x.hi, self.x.lo, self.y, self.dy,
                                     (self.dx_target if self.dx_target is not None else self.x.grad)
                                     if self.needs_dx else None, *self.w[i], eng.partial,
                                     dx_accumulate=self.dx_accumulate, split=eng.split,
                                     alg_k=27 if self.stem else None, fuse_stats=eng.fuse_stats,
                                     dgrad_bn=self._dgrad_bn(eng)) for i in range(2)]

    def _dgrad_bn(self, eng):
        v = self.bn_producer
        if v is None or not eng.fuse_bwd_stats or self.dx_target is not None or self.dx_accumulate:
            return None
        return v.y, v.out.hi, v.mean, v.rstd


class Block:
    def __init__(self):
        self.units = []
        self.ds = None          # downsample Unit
        self.pooled = None      # Act: AvgPool2d(stride) of the block input (stride-2 downsample only)
        self.x = None
        self.out = None


class FullBatchEngine:
    """Runs the per-microbatch gradient + finite-difference regulariser on the sm_100a kernels.

    precision: "split"  -- activations and weights enter the tensor cores as bf16 hi+lo pairs (3 MMAs forward,
                           2 dgrad, 2 wgrad; ~16 mantissa bits per operand): the parity mode;
               "bf16"   -- plain bf16 operands (1 MMA each): the fast mode, cannot resolve the FD perturbation.
    """

    def __init__(self, model, microbatch, precision="split", label_smoothing=0.0, device=None):
        if not isinstance(model, ResNet):
            raise RuntimeError("FullBatchEngine needs a model built by fullbatchtraining_b200.construct_model "
                               "(there is no fallback path)")
        if precision not in ("split", "bf16"):
            raise ValueError(f"unknown precision {precision!r}")
        if not torch.cuda.is_available():
            raise RuntimeError("FullBatchEngine needs a CUDA device (B200); there is no CPU path")
        self.device = torch.device(device or "cuda")
        self.model = model.to(self.device, torch.float32)
        self.mb = int(microbatch)
        self.split = precision == "split"
        self.precision = precision
        self.smoothing = float(label_smoothing)
        self.classes = model.fc.out_features
        self.partial_elems = 0
        # BatchNorm statistics in the conv epilogue (FB_FUSE_STATS=0: statistics pass inside the BatchNorm kernel)
        self.fuse_stats = os.environ.get("FB_FUSE_STATS", "1") == "1"
        # BatchNorm-backward statistics from the epilogue of the dgrad that produces the BatchNorm's upstream gradient
        # (conv2 -> bn1 of every block).  Opt-in: measured -1.5 % on B200 -- the epilogue has to pull the BatchNorm's
        # pre-activation and ReLU mask from HBM, which costs the tensor-core kernel more than the skipped pass saves.
        self.fuse_bwd_stats = os.environ.get("FB_FUSE_BWD_STATS", "0") == "1"
        dev = self.device

        # ---- flat parameter buffers, parameters() order
        self.names, self.offsets, self.shapes = [], {}, {}
        off = 0
        for name, p in self.model.named_parameters():
            self.names.append(name)
            self.offsets[name] = off
            self.shapes[name] = tuple(p.shape)
            off += p.numel()
        self.numel = off
        pad = (-off) % 4
        self.theta = torch.zeros(off + pad, device=dev)[:off]
        self.theta_p = torch.zeros(off + pad, device=dev)[:off]
        self.g = torch.zeros(off + pad, device=dev)[:off]
        self.g2 = torch.zeros(off + pad, device=dev)[:off]
        self.avg = torch.zeros(off + pad, device=dev)[:off]
        with torch.no_grad():
            for name, p in self.model.named_parameters():
                o = self.offsets[name]
                if o % 4 != 0:
                    raise RuntimeError(f"parameter {name} is not 16-byte aligned in the flat buffer")
                self.theta[o:o + p.numel()].copy_(p.reshape(-1))
                p.data = self.theta[o:o + p.numel()].view(p.shape)
        self.scal = torch.zeros(16, device=dev)
        self.cursor = torch.zeros(1, device=dev, dtype=torch.int32)
        self.sq_ws = torch.zeros(1024, device=dev, dtype=torch.float64)
        self.labels_mb = torch.zeros(self.mb, device=dev, dtype=torch.int64)

        # ---- network plan
        n = self.mb
        self.patches = Act(n, 32, 32, 64, self.split, dev, grad=False)
        stem_conv = self.model.stem[0]
        if stem_conv.in_channels != 3 or stem_conv.out_channels != 64:
            raise RuntimeError("the stem kernel expects 3 input channels and 64 output channels")
        self.stem = Unit(self, "stem.0", "stem.1", self.patches, 64, 1, 1, False, needs_dx=False, stem=True)
        self.a0 = Act(n, 32, 32, 64, self.split, dev)
        self.stem.out = self.a0
        self.blocks = []
        cur = self.a0
        max_c = 64
        for s, stage in enumerate(self.model.layers):
            for b, mod in enumerate(stage):
                assert isinstance(mod, ResidualBlock)
                blk = Block()
                blk.x = cur
                pre = f"layers.{s}.{b}"
                x = cur
                pairs = mod.conv_bn_pairs()
                for i, (cn, bnn) in enumerate(pairs):
                    conv = getattr(mod, cn)
                    k, st = conv.kernel_size[0], conv.stride[0]
                    u = Unit(self, f"{pre}.{cn}", f"{pre}.{bnn}", x, conv.out_channels, k, st)
                    u.out = Act(x.n, u.ho, u.wo, conv.out_channels, self.split, dev)
                    if i > 0:
                        u.bn_producer = blk.units[i - 1]
                        blk.units[i - 1].bn_consumer = u
                    blk.units.append(u)
                    x = u.out
                    max_c = max(max_c, conv.out_channels)
                cur.add_grad2()  # shortcut-branch gradient of this block's input
                if mod.downsample is not None:
                    pool, dconv = mod.downsample[0], mod.downsample[1]
                    ps = pool.kernel_size if isinstance(pool.kernel_size, int) else pool.kernel_size[0]
                    src, target = cur, cur.grad2
                    if ps == 2:
                        blk.pooled = Act(cur.n, cur.h // 2, cur.w // 2, cur.c, self.split, dev)
                        src, target = blk.pooled, None
                    elif ps != 1:
                        raise RuntimeError(f"AvgPool2d({ps}) in the shortcut is not supported")
                    blk.ds = Unit(self, f"{pre}.downsample.1", f"{pre}.downsample.2", src, dconv.out_channels, 1, 1,
                                  dx_target=target)
                blk.out = x
                self.blocks.append(blk)
                cur = x
        self.last = cur
        self.partial = torch.zeros(self.partial_elems, device=dev)
        self.bn_ws = torch.zeros(2 * max_c * 1024, device=dev)
        self.head_ws = torch.zeros(n * (cur.c + 32), device=dev)
        self.units = [self.stem] + [u for blk in self.blocks for u in (blk.units + ([blk.ds] if blk.ds else []))]
        for u in self.units:
            u.finish(self)
        self.wprep = []
        for i in range(2):
            entries = [(self.offsets[u.conv_name + ".weight"], 64 if u.stem else u.cout, 3 if u.stem else u.cin,
                        9 if u.stem else u.taps, *u.w[i]) for u in self.units]
            self.wprep.append(ops.WeightPrepTable(entries, dev))
        self._bn_modules = dict(self.model.named_modules())
        self._graphs = {}
        # FB_WGRAD_STREAM=0 disables the side stream (everything in one stream)
        self.wgrad_stream = torch.cuda.Stream(device=dev) if os.environ.get("FB_WGRAD_STREAM", "1") == "1" else None
        self.grad_norms = None
        self.aug_params, self.aug_mean, self.aug_std = None, [0.0, 0.0, 0.0], [1.0, 1.0, 1.0]
        self.norm_offset = 0
        self.bn_passes = 0  # number of train-mode forward passes since the last sync of num_batches_tracked

    # ------------------------------------------------------------------------------------------------------------
    def _view(self, flat, name):
        o = self.offsets[name]
        n = 1
        for d in self.shapes[name]:
            n *= d
        return flat[o:o + n]

    def _bn_buffers(self, bn_name):
        m = self._bn_modules[bn_name]
        return m.running_mean, m.running_var

    def _bn_params(self, u, P):
        return self._view(P, u.bn_name + ".weight"), self._view(P, u.bn_name + ".bias")

    def _bn_forward(self, u, P, out, relu=True, second=None, res=None):
        """fused train-mode BatchNorm of unit `u` (statistics, running-stat EMA, normalise, add, ReLU) -> `out` planes"""
        ga, be = self._bn_params(u, P)
        sec = None
        if second is not None:
            ga2, be2 = self._bn_params(second, P)
            sec = (second.y, second.mean, second.rstd, ga2, be2, *self._bn_buffers(second.bn_name))
        ops.bn_fwd_fused(u.y, u.mean, u.rstd, ga, be, u.P, u.cout, out.hi, out.lo, self.bn_ws,
                         running=self._bn_buffers(u.bn_name), relu=relu, second=sec, res=res, momentum=BN_MOMENTUM,
                         eps=BN_EPS, stats=u.plans[self._pass].stats,
                         stats2=second.plans[self._pass].stats if second is not None else None)

    def _forward(self, P, G, loss_slot, correct_slot):
        self._pass = 0 if P is self.theta else 1
        if self._pass == 1:
            self.wprep[1](P)  # operands of theta' = theta + eps_n*v; those of theta are refreshed once per step
        u = self.stem
        u.plans[self._pass].forward()
        self._bn_forward(u, P, u.out)
        for blk in self.blocks:
            last = len(blk.units) - 1
            for i, u in enumerate(blk.units):
                u.plans[self._pass].forward()
                if i < last:
                    self._bn_forward(u, P, u.out)
            u = blk.units[last]
            if blk.ds is not None:
                d = blk.ds
                if blk.pooled is not None:
                    x = blk.x
                    ops.avgpool2_fwd(x.hi, x.lo, x.n, x.h, x.w, x.c, blk.pooled.hi, blk.pooled.lo)
                d.plans[self._pass].forward()
                self._bn_forward(u, P, blk.out, second=d)
            else:
                self._bn_forward(u, P, blk.out, res=(blk.x.hi, blk.x.lo))
        a = self.last
        ops.head_fwd_bwd(a.hi, a.lo, a.n, a.h * a.w, a.c, self._view(P, "fc.weight"), self._view(P, "fc.bias"),
                         self.labels_mb, self.classes, self.smoothing, self.head_ws, self.scal, loss_slot, correct_slot,
                         self._view(G, "fc.weight"), self._view(G, "fc.bias"), a.grad)

    @torch.no_grad()
    def forward_eval(self, x):
        """Eval-mode forward of `x` [n <= microbatch, 3, 32, 32] fp32 on the CUDA kernels (training.py:343-388 calls
        model(inputs) in eval mode): BatchNorm uses the running statistics (fb_bn_apply), nothing is written to the
        model's buffers or gradients.  Returns the logits [n, classes] (the final 512->10 linear on the pooled features
        runs in torch: the caller needs logits, not the loss, for the test_time_flips softmax sum)."""
        n = x.shape[0]
        assert x.is_cuda and x.dtype == torch.float32 and tuple(x.shape[1:]) == (3, 32, 32) and n <= self.mb
        if not hasattr(self, "_eval_x"):
            self._eval_x = torch.zeros(self.mb, 3, 32, 32, device=self.device)
            self._eval_y = torch.zeros(self.mb, dtype=torch.int64, device=self.device)
        self._eval_x.zero_()
        self._eval_x[:n].copy_(x)  # a short last batch is zero padded: eval-mode BN is per sample
        P = self.theta
        self._pass = 0
        self.wprep[0](P)
        ops.stem_im2col(self._eval_x, self._eval_y, None, None, 0, self.mb, self.patches.hi, self.patches.lo,
                        self.labels_mb)

        def bn_eval(u, out, second=None, res=None):
            def stat(v):
                rm, rv = self._bn_buffers(v.bn_name)
                ga, be = self._bn_params(v, P)
                return rm, torch.rsqrt(rv + BN_EPS), ga, be
            rm, rs, ga, be = stat(u)
            sec = None
            if second is not None:
                rm2, rs2, ga2, be2 = stat(second)
                sec = (second.y, rm2, rs2, ga2, be2)
            ops.bn_apply(u.y, rm, rs, ga, be, u.P, u.cout, out.hi, out.lo, relu=True, second=sec, res=res)

        u = self.stem
        u.plans[0].forward()
        bn_eval(u, u.out)
        for blk in self.blocks:
            last = len(blk.units) - 1
            for i, u in enumerate(blk.units):
                u.plans[0].forward()
                if i < last:
                    bn_eval(u, u.out)
            u = blk.units[last]
            if blk.ds is not None:
                d = blk.ds
                if blk.pooled is not None:
                    xin = blk.x
                    ops.avgpool2_fwd(xin.hi, xin.lo, xin.n, xin.h, xin.w, xin.c, blk.pooled.hi, blk.pooled.lo)
                d.plans[0].forward()
                bn_eval(u, blk.out, second=d)
            else:
                bn_eval(u, blk.out, res=(blk.x.hi, blk.x.lo))
        a = self.last
        pooled = (a.hi.float() + (a.lo.float() if a.lo is not None else 0.0)).view(a.n, a.h * a.w, a.c).mean(dim=1)
        logits = torch.addmm(self._view(P, "fc.bias").view(-1), pooled,
                             self._view(P, "fc.weight").view(self.classes, a.c).t())
        return logits[:n]

    def _unit_backward(self, u, P, G, act, dz_out=None):
        """BN(+ReLU) backward of `u` from the gradient of activation `act` (= grad + grad2), then wgrad and dgrad."""
        ga, _ = self._bn_params(u, P)
        consumer = getattr(u, "bn_consumer", None)  # the unit whose dgrad produced act.grad and its statistics
        stats = consumer.plans[self._pass].dgrad_stats if (consumer is not None and act.grad2 is None) else None
        ops.bn_bwd_fused(act.grad, act.hi, u.y, u.mean, u.rstd, ga, u.P, u.cout, self.bn_ws,
                         self._view(G, u.bn_name + ".weight"), self._view(G, u.bn_name + ".bias"), u.dy, dz_out=dz_out,
                         dA2=act.grad2, stats=stats)
        gw = self._view(G, u.conv_name + ".weight")
        plan = u.plans[self._pass]
        # wgrad (+ its split-K reduction) only feeds the flat gradient: it runs on a side stream, concurrently with the
        # dgrad -> BatchNorm-backward chain of the layers below (fork here, join at the end of the backward pass)
        if self.wgrad_stream is not None:
            main = torch.cuda.current_stream()
            self.wgrad_stream.wait_stream(main)
            with torch.cuda.stream(self.wgrad_stream):
                plan.wgrad(gw, cin_real=3, mode=1) if u.stem else plan.wgrad(gw)
        else:
            plan.wgrad(gw, cin_real=3, mode=1) if u.stem else plan.wgrad(gw)
        if not u.stem:
            plan.dgrad()

    def _backward(self, P, G):
        for blk in reversed(self.blocks):
            out = blk.out
            last = len(blk.units) - 1
            if blk.ds is not None:
                # shortcut branch: its input gradient goes to the block input's second gradient buffer (grad2)
                d = blk.ds
                self._unit_backward(d, P, G, out)
                if blk.pooled is not None:
                    x = blk.x
                    ops.avgpool2_bwd(blk.pooled.grad, x.n, x.h, x.w, x.c, x.grad2, accumulate=False)
                dz_out = None
            else:
                dz_out = blk.x.grad2  # identity shortcut: dz of the last BN is the shortcut gradient
            for i in range(last, -1, -1):
                u = blk.units[i]
                if i == last:
                    self._unit_backward(u, P, G, out, dz_out=dz_out)
                else:
                    self._unit_backward(u, P, G, u.out)
        self._unit_backward(self.stem, P, G, self.a0)
        if self.wgrad_stream is not None:
            torch.cuda.current_stream().wait_stream(self.wgrad_stream)  # join: g is complete, activations are free

    # ------------------------------------------------------------------------------------------------------------
    def _microbatch_ops(self, x_src, labels_src, perm, first, use_cursor, block_strength, eps, accumulate, write_g,
                        mode="full", impl="forward", acc=0.0, batch_clip=None, target="avg"):
        """mode "full": whole per-microbatch recipe; "raw": pass 1 only (training.py:76-83 + :162);
        "reg": regulariser only, self.g already holds the raw gradient of this microbatch (modules.py:211-241).
        impl "forward" | "central" (modules.py:211-241 / :266-300); acc: acc_strength with self.pre as pre_grads;
        batch_clip: per-microbatch L2 clip before the running mean (training.py:166-168); target "avg" | "pre"
        (the acc_strength pre-pass of training.py:128-142 accumulates raw gradients into self.pre)."""
        dst = self.avg if target == "avg" else self.pre
        if x_src.dtype == torch.uint8:  # raw HWC dataset: crop / flip / normalise fused into the im2col
            ops.stem_im2col_u8aug(x_src, labels_src, perm, self.cursor if use_cursor else None, first, self.mb,
                                  self.aug_params, self.aug_mean, self.aug_std, self.patches.hi, self.patches.lo,
                                  self.labels_mb)
        else:
            ops.stem_im2col(x_src, labels_src, perm, self.cursor if use_cursor else None, first, self.mb,
                            self.patches.hi, self.patches.lo, self.labels_mb)
        if mode != "reg":
            self._forward(self.theta, self.g, S_LOSS, S_CORRECT)
            self._backward(self.theta, self.g)
        norms = None
        if target == "avg":
            norms = self.grad_norms[self.norm_offset:] if self.norm_offset else self.grad_norms
        ops.flat_sqnorm(self.g, self.numel, self.sq_ws, self.scal, S_N2, norms, self.cursor)
        regularise = (block_strength != 0 or acc != 0) and mode != "raw"
        fused_mean = dst if (accumulate and batch_clip is None) else None
        if regularise:
            if impl == "forward" and acc == 0:
                ops.fd_perturb(self.theta, self.g, self.numel, block_strength, eps, self.scal, S_N2, S_EPS, self.theta_p)
                self._forward(self.theta_p, self.g2, S_LOSS2, S_CORRECT2)
                self._backward(self.theta_p, self.g2)
                ops.fd_combine(self.g, self.g2, fused_mean, self.numel, self.scal, S_EPS, 0.0, self.cursor, 0,
                               write_g or batch_clip is not None, cf_slot=S_CF)
            else:
                pre = self.pre if acc != 0 else None
                ops.flat_sqnorm_axpby(self.g, pre, block_strength, acc, self.numel, self.sq_ws, self.scal, S_VSQ)
                if impl == "forward":
                    ops.fd_perturb_ex(self.theta, self.g, pre, self.numel, block_strength, acc, eps, 1.0, self.scal,
                                      S_VSQ, S_EPS, self.theta_p)
                    self._forward(self.theta_p, self.g2, S_LOSS2, S_CORRECT2)
                    self._backward(self.theta_p, self.g2)
                    ops.fd_combine(self.g, self.g2, fused_mean, self.numel, self.scal, S_EPS, 0.0, self.cursor, 0,
                                   write_g or batch_clip is not None, cf_slot=S_CF)
                else:
                    for scale, gbuf in ((0.5, self.g2), (-0.5, self._g3())):
                        ops.fd_perturb_ex(self.theta, self.g, pre, self.numel, block_strength, acc, eps, scale,
                                          self.scal, S_VSQ, S_EPS, self.theta_p)
                        self._forward(self.theta_p, gbuf, S_LOSS2, S_CORRECT2)
                        self._backward(self.theta_p, gbuf)
                    ops.fd_combine_ex(self.g, self.g2, self._g3(), fused_mean, self.numel, self.scal, S_EPS, 0.0,
                                      self.cursor, 0, write_g or batch_clip is not None, cf_slot=S_CF)
            if accumulate and batch_clip is not None:
                ops.flat_sqnorm(self.g, self.numel, self.sq_ws, self.scal, S_REGSQ)
                ops.mean_accumulate_clip(self.g, dst, self.numel, self.cursor, 0, self.scal, S_REGSQ, batch_clip,
                                         S_CLIPPED)
        elif accumulate:
            if batch_clip is not None:
                ops.mean_accumulate_clip(self.g, dst, self.numel, self.cursor, 0, self.scal, S_N2, batch_clip, S_CLIPPED)
            else:
                ops.mean_accumulate(self.g, dst, self.numel, self.cursor, 0)
        ops.cursor_add(self.cursor, 1)

    def _g3(self):
        if not hasattr(self, "g3"):
            self.g3 = torch.zeros_like(self.g)
        return self.g3

    def _program(self, x_src, labels_src, perm, first, use_cursor, block_strength, eps, accumulate=True, write_g=False,
                 use_graph=True, mode="full", impl="forward", acc=0.0, batch_clip=None, target="avg"):
        """Returns a callable running one microbatch; captured into a CUDA graph on first use."""
        if acc != 0 or target == "pre":
            if not hasattr(self, "pre"):
                self.pre = torch.zeros_like(self.g)
        if impl == "central":
            self._g3()
        args = (x_src, labels_src, perm, first, use_cursor, block_strength, eps, accumulate, write_g, mode, impl, acc,
                batch_clip, target)
        if not use_graph:
            return lambda: self._microbatch_ops(*args)
        key = (x_src.data_ptr(), labels_src.data_ptr(), None if perm is None else perm.data_ptr(), first, use_cursor,
               float(block_strength), float(eps), accumulate, write_g, mode, impl, float(acc), batch_clip, target,
               self.grad_norms.data_ptr(), self.norm_offset,
               None if self.aug_params is None else self.aug_params.data_ptr(), tuple(self.aug_mean), tuple(self.aug_std))
        if key not in self._graphs:
            state = self._save_state()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up launch (sets kernel attributes, loads modules) outside capture
                self._microbatch_ops(*args)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._microbatch_ops(*args)
            self._restore_state(state)
            self._graphs[key] = (graph, (x_src, labels_src, perm))
        return self._graphs[key][0].replay

    def _save_state(self):
        bufs = [b.clone() for b in self.model.buffers()]
        return dict(avg=self.avg.clone(), scal=self.scal.clone(), cursor=self.cursor.clone(), bufs=bufs,
                    norms=self.grad_norms.clone(), g=self.g.clone(),
                    pre=self.pre.clone() if hasattr(self, "pre") else None)

    def _restore_state(self, st):
        self.avg.copy_(st["avg"])
        self.scal.copy_(st["scal"])
        self.cursor.copy_(st["cursor"])
        self.grad_norms.copy_(st["norms"])
        self.g.copy_(st["g"])
        if st["pre"] is not None:
            self.pre.copy_(st["pre"])
        for b, s in zip(self.model.buffers(), st["bufs"]):
            b.copy_(s)

    # ------------------------------------------------------------------------------------------------------------
    # ---- device-side data pipeline ---------------------------------------------------------
    def set_normalization(self, mean, std):
        """Normalize(mean, std) of config/data/CIFAR10.yaml:8-17 for uint8 datasets."""
        self.aug_mean, self.aug_std = [float(v) for v in mean], [float(v) for v in std]

    def draw_augmentation(self, num_samples, generator=None, crop_padding=4, flip=0.5):
        """One epoch of RandomCrop(32, padding) + RandomHorizontalFlip(flip) draws, kept on the device in a persistent
        int8 [num_samples, 4] buffer (dx, dy, flip, 0) indexed by position in the epoch order.  None disables."""
        if self.aug_params is None or self.aug_params.shape[0] < num_samples:
            self.aug_params = torch.zeros(num_samples, 4, device=self.device, dtype=torch.int8)
            self._graphs.clear()
        n = self.aug_params.shape[0]
        off = torch.randint(0, 2 * crop_padding + 1, (n, 2), device=self.device, generator=generator)
        flips = (torch.rand(n, device=self.device, generator=generator) < flip)
        self.aug_params[:, 0:2] = off.to(torch.int8)
        self.aug_params[:, 2] = flips.to(torch.int8)
        return self.aug_params

    def set_lr(self, lr):
        """correction factor lr/4 of modules.py:214, kept on the device so captured graphs stay valid"""
        self.scal[S_CF] = lr / 4

    def begin_step(self, num_microbatches):
        if self.grad_norms is None or self.grad_norms.numel() < num_microbatches:
            self.grad_norms = torch.zeros(max(num_microbatches, 16), device=self.device)
            self._graphs.clear()
        self.grad_norms.zero_()
        self.avg.zero_()
        self.scal[S_LOSS:S_CORRECT2 + 1] = 0
        self.scal[S_CLIPPED] = 0
        self.cursor.zero_()
        self.wprep[0](self.theta)  # theta is constant during the step: pass-1 operands once, not per microbatch

    def accumulate_resident(self, X, Y, lr, block_strength, eps, first=0, count=None, perm=None, use_graph=True,
                            num_norms=None, norm_offset=0, implementation="forward-differences", acc_strength=0.0,
                            batch_clip=None):
        """Full-batch accumulation over `count` consecutive microbatches of a device-resident dataset
        X [N,3,32,32] fp32, Y [N] int64, starting at sample `first` (optionally through the index tensor `perm`).
        Returns after enqueueing; results: self.avg (running mean), self.grad_norms[:count], loss/correct sums in scal."""
        assert X.is_cuda and X.is_contiguous() and Y.dtype == torch.int64
        if X.dtype == torch.uint8:
            assert tuple(X.shape[1:]) == (32, 32, 3), "uint8 datasets are HWC [N,32,32,3]"
        else:
            assert X.dtype == torch.float32 and tuple(X.shape[1:]) == (3, 32, 32)
        n_avail = (perm.numel() if perm is not None else X.shape[0]) - first
        count = n_avail // self.mb if count is None else count
        impl = IMPLEMENTATIONS[implementation]
        self.begin_step(num_norms or count)
        self.norm_offset = int(norm_offset)
        self.set_lr(lr)
        if acc_strength != 0:
            # training.py:128-142: extra sweep for the mean raw gradient (pre_grads); single GPU only for now
            self.pre = torch.zeros_like(self.g) if not hasattr(self, "pre") else self.pre.zero_()
            pre_run = self._program(X, Y, perm, first, True, 0.0, eps, use_graph=use_graph, mode="raw", impl=impl,
                                    batch_clip=batch_clip, target="pre")
            for _ in range(count):
                pre_run()
            self.bn_passes += count
            self.scal[S_LOSS:S_CORRECT2 + 1] = 0
            self.scal[S_CLIPPED] = 0
            self.cursor.zero_()
        run = self._program(X, Y, perm, first, True, block_strength, eps, use_graph=use_graph, impl=impl,
                            acc=acc_strength, batch_clip=batch_clip)
        for _ in range(count):
            run()
        regularised = block_strength != 0 or acc_strength != 0
        self.bn_passes += count * ((3 if impl == "central" else 2) if regularised else 1)
        return count

    def _stages(self):
        if not hasattr(self, "_x_stage"):
            self._x_stage = [torch.zeros(self.mb, 3, 32, 32, device=self.device) for _ in range(2)]
            self._y_stage = [torch.zeros(self.mb, device=self.device, dtype=torch.int64) for _ in range(2)]
            self._stage_free = [torch.cuda.Event() for _ in range(2)]
            self._copy_stream = torch.cuda.Stream()
        return self._x_stage, self._y_stage

    def accumulate_stream(self, loader, lr, block_strength, eps, num_microbatches, use_graph=True, norm_offset=0):
        """Full-batch accumulation over blocks (inputs [B,3,32,32], labels [B]) coming from a host-side iterable (the
        reference's DataLoader protocol, training.py:148-152): every block is split into microbatches
        (torch.chunk, training.py:155-156), copied host->device on a copy stream into one of two staging buffers and
        consumed by a captured graph, so the copy of microbatch k+1 overlaps the compute of microbatch k."""
        xs, ys = self._stages()
        self.begin_step(num_microbatches)
        self.norm_offset = int(norm_offset)
        self.set_lr(lr)
        runs = [self._program(xs[i], ys[i], None, 0, False, block_strength, eps, use_graph=use_graph) for i in range(2)]
        main = torch.cuda.current_stream()
        k = 0
        h2d = 0
        for inputs, labels in loader:
            chunks = max(labels.shape[0] // self.mb, 1)
            for xc, yc in zip(torch.chunk(inputs, chunks, dim=0), torch.chunk(labels, chunks, dim=0)):
                if xc.shape[0] != self.mb:
                    raise RuntimeError(f"microbatch of {xc.shape[0]} samples, engine built for {self.mb} (drop_last?)")
                i = k & 1
                with torch.cuda.stream(self._copy_stream):
                    self._copy_stream.wait_event(self._stage_free[i])
                    xs[i].copy_(xc, non_blocking=True)
                    ys[i].copy_(yc, non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(self._copy_stream)
                main.wait_event(ready)
                runs[i]()
                self._stage_free[i].record(main)
                h2d += xc.numel() * xc.element_size() + yc.numel() * yc.element_size()
                k += 1
        self.bn_passes += k * (2 if block_strength != 0 else 1)
        self.h2d_bytes = h2d
        return k

    # ---- GradRegularizer / _compute_batched_gradient protocol ------------------------------------------------------
    def microbatch_gradient(self, inputs, labels):
        """training.py:76-83 on the device for one microbatch: fills self.g (raw gradient), returns (loss, correct)."""
        xs, ys = self._stages()
        xs[0].copy_(inputs)
        ys[0].copy_(labels)
        self.begin_step(1)
        self._program(xs[0], ys[0], None, 0, False, 0.0, 0.0, accumulate=False, mode="raw")()
        self.bn_passes += 1
        return self.scal[S_LOSS], self.scal[S_CORRECT]

    def regularize(self, inputs, labels, lr, block_strength, eps, implementation="forward-differences",
                   acc_strength=0.0):
        """modules.py:211-300: self.g (raw gradient of this microbatch) <- regularised gradient, in place
        (acc_strength uses self.pre, see load_pre)."""
        xs, ys = self._stages()
        xs[0].copy_(inputs)
        ys[0].copy_(labels)
        if self.grad_norms is None:
            self.begin_step(1)
        self.cursor.zero_()
        self.set_lr(lr)
        self._pass = 1
        impl = IMPLEMENTATIONS[implementation]
        self._program(xs[0], ys[0], None, 0, False, block_strength, eps, accumulate=False, write_g=True, mode="reg",
                      impl=impl, acc=acc_strength)()
        self.bn_passes += 2 if impl == "central" else 1

    def load_pre(self, pre_grads):
        """pre_grads (list shaped like model.parameters(), training.py:128-142) -> flat self.pre"""
        if not hasattr(self, "pre"):
            self.pre = torch.zeros_like(self.g)
        for name, t in zip(self.names, pre_grads):
            self._view(self.pre, name).copy_(t.reshape(-1))

    def load_grads(self, grads):
        for name, gt in zip(self.names, grads):
            v = self._view(self.g, name)
            if gt.data_ptr() != v.data_ptr():
                v.copy_(gt.reshape(-1))

    def store_grads(self, grads):
        for name, gt in zip(self.names, grads):
            v = self._view(self.g, name)
            if gt.data_ptr() != v.data_ptr():
                gt.copy_(v.view(gt.shape))

    # ---- multi-GPU: one all-reduce of the flat buffer per full-batch pass (training/utils.py:31-41) -----------------
    def all_reduce_mean(self, local_count, global_count):
        """avg holds the running mean over this rank's `local_count` microbatches; after this call every rank holds the
        mean over all `global_count` microbatches (exactly the single-process result up to fp32 summation order).
        The reference's own multi-process weighting (training.py:168 with num_machines > 1) is not a mean and is
        deliberately not reproduced."""
        import torch.distributed as dist

        ops.flat_scale(self.avg, self.numel, float(local_count) / float(global_count))
        dist.all_reduce(self.avg, op=dist.ReduceOp.SUM)
        pack = torch.cat([self.scal[S_LOSS:S_CORRECT + 1], self.grad_norms])
        dist.all_reduce(pack, op=dist.ReduceOp.SUM)
        self.scal[S_LOSS:S_CORRECT + 1] = pack[:2]
        self.grad_norms.copy_(pack[2:])

    def results(self, count):
        """Host read of the step scalars (one synchronisation): mean loss, correct count, grad_norms."""
        s = self.scal.tolist()
        return dict(loss=s[S_LOSS] / max(count, 1), correct=s[S_CORRECT], loss_sum=s[S_LOSS],
                    grad_norms=self.grad_norms[:count].clone(), clipped_batches=int(s[S_CLIPPED]))

    def sync_bn_counters(self):
        """num_batches_tracked += number of train-mode passes (2 per microbatch with the regulariser)."""
        if self.bn_passes:
            for m in self.model.modules():
                if isinstance(m, torch.nn.BatchNorm2d):
                    m.num_batches_tracked += self.bn_passes
            self.bn_passes = 0

    def grads_list(self, flat):
        """Views of a flat buffer shaped like model.parameters()."""
        return [self._view(flat, n).view(self.shapes[n]) for n in self.names]
