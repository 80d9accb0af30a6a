"""Executor of the full-batch gradient-regularised step on one B200, G microbatches per launch.

Replaces the device work of ``_accumulate_full_gradient`` (reference fullbatch/training/training.py:121-185) and
``GradRegularizer._forward_differences`` (fullbatch/models/modules.py:211-241).  The reference walks its microbatches
one by one; here ONE replay of a captured CUDA graph serves a GROUP LAUNCH of up to G consecutive microbatches ("groups"
of the C ABI): every kernel of the network takes the group dimension, so a ResNet-18 step over 390 microbatches of 128
images is 49 replays of ~190 large launches instead of 390 replays of 262 small ones.  Per group launch:

    im2col of the ng microbatches (stem)                                          [fb_stem_im2col]
    pass 1: forward / loss / backward at theta (weights shared by the groups) -> g[0..ng)     [tcgen05 convs + layer kernels]
    n2_k = |g_k|^2 -> grad_norms[k]; eps_k = eps / (bs*sqrt(n2_k))                [fb_flat_sqnorm]
    pass 2: forward / backward of microbatch k at ITS point theta + eps_k*bs*g_k -> g2[0..ng)
            (conv operands of the perturbed points come straight from fb_weight_prep_multi; BatchNorm / fc parameters from
            fb_perturb_ranges; the weight tile of a GEMM tile is picked by the tile's group)
    for k in loader order: g_reg = g_k + (lr/4)(g2_k - g_k)/eps_k ; avg += (g_reg - avg)/(k+1)          [fb_fd_combine]
    BatchNorm running statistics: EMA in the reference's order (p1_k, p2_k, p1_{k+1}, ...)              [fb_bn_ema_multi]

BatchNorm statistics, loss, gradient norm and weight gradient of a microbatch never mix with another microbatch's, and
every reduction order is a function of one microbatch's problem only: results are bit-identical for every G.

Persistent state (allocated once, kernels never allocate):
  theta                     : flat fp32 parameters in model.parameters() order (= training/utils.py:34 order); the model's
                              parameters are re-pointed to views of it so optimizers update it in place
  theta_p, g_n, g2_n        : [G][stride] perturbed non-conv parameters / per-microbatch gradients (conv weights in the
                              kernels' native [co][tap][ci] layout)
  avg_n -> avg              : running mean in native layout; converted once per step to `avg` (reference layout, the
                              storage of param.grad)
  per conv: bf16 hi/lo GEMM operands of theta (one set) and of the G perturbed points
  per layer: fp32 conv output, bf16 hi/lo activation planes, fp32 activation gradient, bf16 output gradient, [G*mb] images

theta itself is never modified by the regulariser, so the reference's "restore from a clone" (modules.py:237-238) is
exact by construction.
"""
import os
import weakref

import torch

from . import lib as L
from . import ops
from .models import ResNet, ResidualBlock

# scal layout: single slots, then arrays of FB_MAX_GROUPS per-group slots
S_LOSS, S_CORRECT, S_CF, S_CLIPPED = 0, 1, 2, 3
S_GNORM, S_PNORM = 7, 8  # used by optim.FlatSGD
S_N2G, S_EPSG, S_LOSSG, S_CORRG, S_LOSS2G, S_CORR2G, S_VSQG, S_REGSQG = (16 * i for i in range(1, 9))
SCAL_SLOTS = 16 * 9

# hyp.grad_reg.implementation -> device recipe.  `forward-differences-legacy` (modules.py:243-264) perturbs along g
# instead of bs*g and scales the correction by bs: the same update up to rounding, so it shares the forward recipe.
IMPLEMENTATIONS = {"finite_diff": "forward", "forward-differences": "forward", "forward-differences-legacy": "forward",
                   "central-differences": "central"}
BN_EPS = 1e-5
BN_MOMENTUM = 0.1
MAX_PASSES = 3  # train-mode forward passes of one group launch: pass 1 + forward FD (2) or central FD (3)


def default_groups(microbatch):
    """Microbatches per launch: ~1024 images (enough 128-pixel tiles for several waves of 148 SMs on the 4x4 stage), at
    most 8 -- beyond that a second lane of 8 is worth more than a wider launch (ResNet-152 mb 32: 3.69k vs 3.58k img/s)."""
    env = os.environ.get("FB_GROUPS")
    if env:
        return max(1, min(L.FB_MAX_GROUPS, int(env)))
    return max(1, min(8, L.FB_MAX_GROUPS, 1024 // max(int(microbatch), 1)))


def launch_sizes(count, groups):
    """Microbatches per group launch for a pass over `count` microbatches: as few launches as `groups` allows, sizes
    equal up to one (49 -> 7 x 7 rather than 6 x 8 + 1: a launch of one group costs about as much as one of three).
    Results do not depend on the sizes (every reduction order is a function of ONE group's problem)."""
    if count <= 0:
        return []
    n = -(-count // groups)
    q, r = divmod(count, n)
    return [q + 1] * r + [q] * (n - r)


class Act:
    """An activation tensor [n,h,w,c]: bf16 hi/lo planes (+ fp32 gradient buffer when it needs one)."""

    def __init__(self, n, h, w, c, split, device, grad=True):
        self.n, self.h, self.w, self.c = n, h, w, c
        self.hi = torch.zeros(n, h, w, c, device=device, dtype=torch.bfloat16)
        self.lo = torch.zeros(n, h, w, c, device=device, dtype=torch.bfloat16) if split else None
        self.grad = torch.zeros(n, h, w, c, device=device, dtype=torch.float32) if grad else None
        # ReLU mask as a bit plane (written by the BatchNorm that produces this activation, read by its backward)
        self.mask = torch.zeros(n * h * w * c // 8, device=device, dtype=torch.uint8) if grad else None
        # second addend of the gradient (shortcut branch of the consuming block); consumers read grad + grad2, so no
        # GEMM epilogue ever has to read-modify-write
        self.grad2 = None

    def add_grad2(self):
        self.grad2 = torch.zeros_like(self.grad)
        return self.grad2


class Unit:
    """conv (bias-free) + BatchNorm: buffers, descriptors and parameter offsets."""

    def __init__(self, eng, conv_name, bn_name, x, cout, k, stride, needs_dx=True, stem=False, dx_target=None):
        dev, split, G = eng.device, eng.split, eng.G
        shared = eng.root.unit_by_name.get(conv_name) if eng.root is not eng else None
        self.conv_name, self.bn_name, self.x, self.stem = conv_name, bn_name, x, stem
        self.cin, self.cout, self.k, self.stride = x.c, cout, k, stride
        n, h, w = x.n, x.h, x.w
        self.ho, self.wo = h // stride, w // stride
        self.Pg = eng.mb * self.ho * self.wo  # pixels per group
        self.y = torch.zeros(n, self.ho, self.wo, cout, device=dev)
        self.dy = torch.zeros(n, self.ho, self.wo, cout, device=dev, dtype=torch.bfloat16)
        self.mean = torch.zeros(G, cout, device=dev)
        self.rstd = torch.zeros(G, cout, device=dev)
        self.bn_batch = torch.zeros(MAX_PASSES, G, 2, cout, device=dev)  # batch mean / unbiased variance per pass, group
        taps = k * k
        self.taps = taps
        bf = dict(device=dev, dtype=torch.bfloat16)
        # bf16 GEMM operands: [0] from theta (refreshed once per step), [1] of the G perturbed points (every launch)
        self.w = []
        for si, rows in enumerate((1, G)):
            if si == 0 and shared is not None:  # the operands of theta are the same for every lane
                self.w.append(shared.w[0])
                continue
            self.w.append((torch.zeros(rows * cout, taps * x.c, **bf),
                           torch.zeros(rows * cout, taps * x.c, **bf) if split else None,
                           torch.zeros(rows * x.c, taps * cout, **bf) if needs_dx else None,
                           torch.zeros(rows * x.c, taps * cout, **bf) if (needs_dx and split) else None))
        self.w_offset = eng.offsets[conv_name + ".weight"]
        self.gamma_off = eng.offsets[bn_name + ".weight"]
        self.beta_off = eng.offsets[bn_name + ".bias"]
        dx = (dx_target if dx_target is not None else x.grad) if needs_dx else None
        self.plan = ops.Conv2dPlan(eng.mb, G, h, w, x.c, cout, k, stride, x.hi, x.lo, self.y, self.dy, dx, self.w,
                                   self.w_offset, split=split, alg_k=27 if stem else None,
                                   grad_cols=27 if stem else None, bn=(self.mean, self.rstd, BN_EPS),
                                   policy_groups=eng.policy_groups, fwd_hi_only=eng.precision == "split_w")
        self.out = None
        eng.unit_by_name[conv_name] = self

    def bn_batch_ptr(self, pass_idx):
        return self.bn_batch.data_ptr() + pass_idx * self.bn_batch.stride(0) * 4


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
    groups:    microbatches per launch (1..16; default ~1024 images); a pure performance knob, results do not depend on it.
    lanes:     independent sets of activation / gradient buffers (default 2 if they fit): consecutive group launches
               alternate between the lanes on separate streams, so the tensor-bound kernels of one launch overlap the
               bandwidth-bound kernels of the other; the order-sensitive tail of every launch (finite-difference
               combine, running mean, running-statistics EMA) runs on the main stream in loader order, so the result is
               bit-identical to one lane.
    """

    def __init__(self, model, microbatch, precision="split", label_smoothing=0.0, device=None, groups=None,
                 policy_groups=None, lanes=None, _parent=None):
        if not isinstance(model, ResNet):
            raise RuntimeError("FullBatchEngine needs a model built by fullbatchtraining_b200.construct_model "
                               "(there is no fallback path)")
        if precision not in ("split", "bf16", "split_w"):
            raise ValueError(f"unknown precision {precision!r}")
        if not torch.cuda.is_available():
            raise RuntimeError("FullBatchEngine needs a CUDA device (B200); there is no CPU path")
        self.device = torch.device(device or "cuda")
        # lane 0 ("root") owns everything the lanes share; further lanes only hold a weak reference to it, so an engine
        # is released by reference counting (its buffers are tens of GB: it must not wait for the cycle collector)
        self._parent = weakref.ref(_parent) if _parent is not None else None
        self._children = []
        self.unit_by_name = {}
        if _parent is None:
            torch.cuda.synchronize(self.device)
            mem0 = torch.cuda.memory_allocated(self.device)
        self.model = model if _parent is not None else model.to(self.device, torch.float32)
        self.mb = int(microbatch)
        self.G = int(groups) if groups else default_groups(self.mb)
        if not 1 <= self.G <= L.FB_MAX_GROUPS:
            raise ValueError(f"groups must be in 1..{L.FB_MAX_GROUPS}")
        # launches are tuned for `policy_groups` microbatches (default ops.POLICY_GROUPS = 8; the stochastic branch, which
        # only ever launches one, passes 1).  Fixed per engine, independent of G: results do not depend on G.
        self.policy_groups = int(policy_groups or ops.POLICY_GROUPS)
        self.split = precision in ("split", "split_w")  # "split_w": numerics ablation, forward reads x_hi only
        self.precision = precision
        self.smoothing = float(label_smoothing)
        self.classes = model.fc.out_features
        dev, G = self.device, self.G

        # ---- flat parameter buffers, parameters() order
        if _parent is not None:  # a further lane: parameters, running mean and the boundary buffers are lane 0's
            r = _parent
            self.names, self.offsets, self.shapes, self.numel, self.stride = r.names, r.offsets, r.shapes, r.numel, r.stride
            self.theta, self.avg_n, self.avg, self.g = r.theta, r.avg_n, r.avg, r.g
            off = self.numel
        else:
            self.names, self.offsets, self.shapes = [], {}, {}
            off = 0
            for name, p in self.model.named_parameters():
                self.names.append(name)
                self.offsets[name] = off
                self.shapes[name] = tuple(p.shape)
                off += p.numel()
            self.numel = off
            self.stride = (off + 63) // 64 * 64  # elements between the flat buffers of consecutive groups
            self.theta = torch.zeros(self.stride, device=dev)[:off]
            self.avg_n = torch.zeros(self.stride, device=dev)[:off]
            self.avg = torch.zeros(self.stride, device=dev)[:off]  # running mean in the reference's layout (param.grad)
            self.g = torch.zeros(self.stride, device=dev)[:off]    # gradient of ONE microbatch in the reference's layout
            with torch.no_grad():
                for name, p in self.model.named_parameters():
                    o = self.offsets[name]
                    if o % 4 != 0:
                        raise RuntimeError(f"parameter {name} is not 16-byte aligned in the flat buffer")
                    self.theta[o:o + p.numel()].copy_(p.reshape(-1))
                    p.data = self.theta[o:o + p.numel()].view(p.shape)
        self.theta_p = torch.zeros(G, self.stride, device=dev)
        self.g_n = torch.zeros(G, self.stride, device=dev)    # per-microbatch gradients, native layout
        self.g2_n = torch.zeros(G, self.stride, device=dev)
        self.scal = torch.zeros(SCAL_SLOTS, device=dev)
        self.cursor = torch.zeros(1, device=dev, dtype=torch.int32)
        self.sq_ws = torch.zeros(1024 * G, device=dev, dtype=torch.float64)
        self.labels_mb = torch.zeros(G * self.mb, device=dev, dtype=torch.int64)

        # ---- network plan
        n = G * self.mb
        self.patches = Act(n, 32, 32, 64, self.split, dev, grad=False)
        stem_conv = self.model.stem[0]
        if stem_conv.in_channels != 3 or stem_conv.out_channels != 64:
            raise RuntimeError("the stem kernel expects 3 input channels and 64 output channels")
        self.stem = Unit(self, "stem.0", "stem.1", self.patches, 64, 1, 1, needs_dx=False, stem=True)
        self.a0 = Act(n, 32, 32, 64, self.split, dev)
        self.stem.out = self.a0
        self.blocks = []
        cur = self.a0
        for s, stage in enumerate(self.model.layers):
            for b, mod in enumerate(stage):
                assert isinstance(mod, ResidualBlock)
                blk = Block()
                blk.x = cur
                pre = f"layers.{s}.{b}"
                x = cur
                for cn, bnn in mod.conv_bn_pairs():
                    conv = getattr(mod, cn)
                    k, st = conv.kernel_size[0], conv.stride[0]
                    u = Unit(self, f"{pre}.{cn}", f"{pre}.{bnn}", x, conv.out_channels, k, st)
                    u.out = Act(x.n, u.ho, u.wo, conv.out_channels, self.split, dev)
                    blk.units.append(u)
                    x = u.out
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
        self.units = [self.stem] + [u for blk in self.blocks for u in (blk.units + ([blk.ds] if blk.ds else []))]

        # ---- workspaces and device tables
        need = sum(u.plan.partial_elems() for u in self.units)
        self.partial = torch.zeros(max(need, 4), device=dev)
        entries, o = [], 0
        for u in self.units:
            pe = u.plan.partial_elems()
            if pe:
                entries.append(u.plan.bind_partial(self.partial[o:o + pe]))
                o += pe
        self.reduce = ops.ReduceTable(entries, dev) if entries else None
        self.bn_ws = torch.zeros(max(ops.bn_bwd_ws_floats(u.Pg, u.cout, G, self.policy_groups) for u in self.units),
                                 device=dev)
        self.head_ws = torch.zeros(ops.head_ws_floats(n, cur.c), device=dev)
        self.wprep = []
        for i in range(2):
            ent = [(u.w_offset, 64 if u.stem else u.cout, 3 if u.stem else u.cin, 9 if u.stem else u.taps, *u.w[i])
                   for u in self.units]
            self.wprep.append(ops.WeightPrepTable(ent, dev, per_group=(i == 1)))
        # parameters that are not conv weights (BatchNorm weight / bias, fc): perturbed into theta_p by fb_perturb_ranges
        conv_w = {u.conv_name + ".weight" for u in self.units}
        ranges, t0 = [], 0
        for name in self.names:
            if name in conv_w:
                continue
            cnt = 1
            for d in self.shapes[name]:
                cnt *= d
            if ranges and ranges[-1][0] + ranges[-1][1] == self.offsets[name]:
                ranges[-1][1] += cnt
            else:
                ranges.append([self.offsets[name], cnt, t0])
            t0 += cnt
        pos = 0
        for r in ranges:  # third column: index of the first thread of the range
            r[2] = pos
            pos += r[1]
        self.ranges = torch.tensor(ranges, dtype=torch.int64, device=dev)
        self.n_ranges, self.ranges_total = len(ranges), pos
        # 3x3 convs whose native gradient layout [co][tap][ci] differs from OIHW: (offset, cout, cin, taps, first block)
        table, blk0 = [], 0
        for u in self.units:
            if u.taps > 1 and not u.stem:
                table.append([u.w_offset, u.cout, u.cin, u.taps, blk0])
                blk0 += u.cout
        self.relayout_table = torch.tensor(table, dtype=torch.int64, device=dev) if table else None
        self.relayout_entries, self.relayout_blocks = len(table), blk0
        self._bn_modules = dict(self.model.named_modules())
        self.ema = ops.BnEmaTable([(self._bn_modules[u.bn_name].running_mean, self._bn_modules[u.bn_name].running_var,
                                    u.bn_batch, u.bn_batch.stride(0), u.cout) for u in self.units], dev)
        self._graphs = {}
        self.l2_order = os.environ.get("FB_L2_ORDER", "1") == "1"
        # ReLU masks as bit planes (1/16 of the bytes of the bf16 plane the backward would otherwise read twice)
        self.bit_masks = os.environ.get("FB_BIT_MASKS", "1") == "1"
        # projection-shortcut blocks: the main branch's last BatchNorm backward leaves dz in out.grad for the shortcut's
        self.dz_hand_over = os.environ.get("FB_DZ_HAND_OVER", "1") == "1"
        # FB_WGRAD_STREAM=1: wgrad on a side stream, concurrently with the dgrad -> BatchNorm-backward chain below it
        self.wgrad_mode = int(os.environ.get("FB_WGRAD_STREAM", "0"))
        self.wgrad_stream = torch.cuda.Stream(device=dev) if self.wgrad_mode else None
        self.grad_norms = None
        self.aug_params, self.aug_mean, self.aug_std = None, [0.0, 0.0, 0.0], [1.0, 1.0, 1.0]
        self.norm_offset = 0
        self.bn_passes = 0  # number of train-mode forward passes since the last sync of num_batches_tracked
        self.h2d_bytes = 0
        self.stream = torch.cuda.Stream(device=dev)       # this lane's compute stream
        self.commit_done = torch.cuda.Event()             # its previous launch has been combined: g / g2 are free again
        if _parent is None:
            # further lanes, if their buffers fit next to lane 0's (FB_LANES / lanes= overrides; 1 = no concurrency)
            # default: 2; 3 where a launch is small (< 1,024 images: ResNet-152 at microbatch 32 gains another 3.5 %)
            want = int(lanes or os.environ.get("FB_LANES", "2" if self.G * self.mb >= 1024 else "3"))
            torch.cuda.synchronize(dev)
            footprint = torch.cuda.memory_allocated(dev) - mem0
            for _ in range(1, max(want, 1)):
                free = torch.cuda.mem_get_info(dev)[0]
                if footprint * 1.15 > free * 0.8:
                    break
                self._children.append(FullBatchEngine(model, microbatch, precision, label_smoothing, device, self.G,
                                                      policy_groups, _parent=self))

    @property
    def root(self):
        return self if self._parent is None else self._parent()

    @property
    def lanes(self):
        return [self] + self._children

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

    def to_native(self, src, dst):
        """flat buffer in the reference's order (OIHW conv weights) -> the kernels' native gradient layout"""
        ops.flat_relayout(src, dst, self.numel, self.relayout_table, self.relayout_entries, self.relayout_blocks, True)

    def from_native(self, src, dst):
        ops.flat_relayout(src, dst, self.numel, self.relayout_table, self.relayout_entries, self.relayout_blocks, False)

    # ---- one pass over the network for ng groups -----------------------------------------------------------------
    def _bn_forward(self, u, ng, P, pstride, out, relu=True, second=None, res=None):
        """train-mode BatchNorm of unit `u` with the statistics its convolution left in u.mean / u.rstd -> `out`"""
        base = P.data_ptr()
        sec = None
        if second is not None:
            sec = (second.y, second.mean, second.rstd, base + 4 * second.gamma_off, base + 4 * second.beta_off)
        # l2_order: walk back to front = start on what the convolution wrote last (still in L2) and end where the next
        # convolution starts
        ops.bn_apply(u.y, u.mean, u.rstd, base + 4 * u.gamma_off, base + 4 * u.beta_off, u.Pg, u.cout, out.hi, out.lo,
                     relu=relu, second=sec, res=res, ng=ng, param_gstride=pstride, reverse=self.l2_order,
                     mask_out=out.mask if self.bit_masks else None)

    def _forward(self, ng, wset, P, pstride, Gbuf, loss_base, correct_base, pass_idx):
        """P: flat parameters (theta: pstride 0, shared; theta_p: one row per group); Gbuf: [G][stride] gradients"""
        u = self.stem
        u.plan.forward(ng, wset, u.bn_batch_ptr(pass_idx))
        self._bn_forward(u, ng, P, pstride, u.out)
        for blk in self.blocks:
            last = len(blk.units) - 1
            for i, u in enumerate(blk.units):
                u.plan.forward(ng, wset, u.bn_batch_ptr(pass_idx))
                if i < last:
                    self._bn_forward(u, ng, P, pstride, u.out)
            u = blk.units[last]
            if blk.ds is not None:
                d = blk.ds
                if blk.pooled is not None:
                    x = blk.x
                    ops.avgpool2_fwd(x.hi, x.lo, ng * self.mb, x.h, x.w, x.c, blk.pooled.hi, blk.pooled.lo)
                d.plan.forward(ng, wset, d.bn_batch_ptr(pass_idx))
                self._bn_forward(u, ng, P, pstride, blk.out, second=d)
            else:
                self._bn_forward(u, ng, P, pstride, blk.out, res=(blk.x.hi, blk.x.lo))
        a = self.last
        pb, gb = P.data_ptr(), Gbuf.data_ptr()
        fw, fb = 4 * self.offsets["fc.weight"], 4 * self.offsets["fc.bias"]
        ops.head_fwd_bwd(a.hi, a.lo, self.mb, a.h * a.w, a.c, pb + fw, pb + fb, self.labels_mb, self.classes,
                         self.smoothing, self.head_ws, self.scal, loss_base, correct_base, gb + fw, gb + fb, a.grad,
                         ng=ng, param_gstride=pstride, grad_gstride=self.stride)

    def _unit_backward(self, u, ng, wset, P, pstride, Gbuf, act, dz_out=None, premasked=False):
        """BN(+ReLU) backward of `u` from the gradient of activation `act` (= grad + grad2), then wgrad and dgrad.
        premasked: act.grad already holds dz = mask * (grad + grad2) (left there by the BatchNorm backward of the other
        branch that ends in `act`)."""
        pb, gb = P.data_ptr(), Gbuf.data_ptr()
        # l2_order: the reduce pass starts where the producer of the upstream gradient ended, the apply pass walks the
        # other way, the dgrad starts where the apply ended -- the direction alternates from layer to layer so that each
        # kernel begins on the ~100 MB its predecessor left in L2
        rev = self._rev and self.l2_order
        self._rev = not self._rev
        if premasked:
            ops.bn_bwd(act.grad, None, u.y, u.mean, u.rstd, pb + 4 * u.gamma_off, u.Pg, u.cout, self.bn_ws,
                       gb + 4 * u.gamma_off, gb + 4 * u.beta_off, u.dy, dz_out=None, dA2=None, ng=ng,
                       param_gstride=pstride, grad_gstride=self.stride, reverse=rev, mask_bits=None,
                       policy_groups=self.policy_groups)
        else:
            ops.bn_bwd(act.grad, act.hi, u.y, u.mean, u.rstd, pb + 4 * u.gamma_off, u.Pg, u.cout, self.bn_ws,
                       gb + 4 * u.gamma_off, gb + 4 * u.beta_off, u.dy, dz_out=dz_out, dA2=act.grad2, ng=ng,
                       param_gstride=pstride, grad_gstride=self.stride, reverse=rev,
                       mask_bits=act.mask if self.bit_masks else None, policy_groups=self.policy_groups)
        # wgrad only feeds the flat gradient.  wgrad_mode 1 / 2: on a side stream, forked before / after the dgrad of the
        # same layer (2: the tensor-bound wgrad then runs next to the bandwidth-bound BatchNorm backward of the layer
        # below instead of next to its own dgrad); joined at the end of the backward pass
        if self.wgrad_stream is not None and self.wgrad_mode == 1:
            self._wgrad_side(u, ng, Gbuf)
        elif self.wgrad_stream is None:
            u.plan.wgrad(ng, Gbuf, self.stride)
        if not u.stem:
            u.plan.dgrad(ng, wset, reverse=rev)
        if self.wgrad_stream is not None and self.wgrad_mode == 2:
            self._wgrad_side(u, ng, Gbuf)

    def _shortcut_backward(self, blk, ng, wset, P, pstride, Gbuf, premasked):
        """Projection shortcut: its input gradient goes to the block input's second gradient buffer (grad2)."""
        self._unit_backward(blk.ds, ng, wset, P, pstride, Gbuf, blk.out, premasked=premasked)
        if blk.pooled is not None:
            x = blk.x
            ops.avgpool2_bwd(blk.pooled.grad, ng * self.mb, x.h, x.w, x.c, x.grad2, accumulate=False)

    def _wgrad_side(self, u, ng, Gbuf):
        main = torch.cuda.current_stream()
        self.wgrad_stream.wait_stream(main)
        with torch.cuda.stream(self.wgrad_stream):
            u.plan.wgrad(ng, Gbuf, self.stride)

    def _backward(self, ng, wset, P, pstride, Gbuf):
        self._rev = True  # the head wrote the first upstream gradient front to back
        for blk in reversed(self.blocks):
            out = blk.out
            last = len(blk.units) - 1
            # dz = mask * (grad + grad2) of the block output is what both branches that end in it start from.  Identity
            # shortcut: the last BatchNorm's backward writes it to the block input's grad2.  Projection shortcut: it
            # writes it over out.grad in place (element-wise, by the thread that read the element), and the shortcut's
            # BatchNorm backward then reads ONE premasked tensor instead of two addends and the mask, twice
            hand_over = blk.ds is not None and self.dz_hand_over
            dz_out = blk.x.grad2 if blk.ds is None else (out.grad if hand_over else None)
            if blk.ds is not None and not hand_over:
                self._shortcut_backward(blk, ng, wset, P, pstride, Gbuf, premasked=False)
            for i in range(last, -1, -1):
                u = blk.units[i]
                if i == last:
                    self._unit_backward(u, ng, wset, P, pstride, Gbuf, out, dz_out=dz_out)
                    if hand_over:
                        self._shortcut_backward(blk, ng, wset, P, pstride, Gbuf, premasked=True)
                else:
                    self._unit_backward(u, ng, wset, P, pstride, Gbuf, u.out)
        self._unit_backward(self.stem, ng, wset, P, pstride, Gbuf, self.a0)
        if self.wgrad_stream is not None:
            torch.cuda.current_stream().wait_stream(self.wgrad_stream)  # join
        if self.reduce is not None:
            self.reduce(ng, Gbuf, self.stride)  # split-K partials of all layers -> flat gradient, fixed order

    @torch.no_grad()
    def forward_eval(self, x):
        """Eval-mode forward of `x` [n <= G*microbatch, 3, 32, 32] fp32 on the CUDA kernels (training.py:343-388 calls
        model(inputs) in eval mode): BatchNorm uses the running statistics (fb_bn_apply), nothing is written to the
        model's buffers or gradients.  Returns the logits [n, classes] (the final 512->10 linear on the pooled features
        runs in torch: the caller needs logits, not the loss, for the test_time_flips softmax sum)."""
        n = x.shape[0]
        assert x.is_cuda and x.dtype == torch.float32 and tuple(x.shape[1:]) == (3, 32, 32) and n <= self.G * self.mb
        ng = -(-n // self.mb)
        if not hasattr(self, "_eval_x"):
            self._eval_x = torch.zeros(self.G * self.mb, 3, 32, 32, device=self.device)
            self._eval_y = torch.zeros(self.G * self.mb, dtype=torch.int64, device=self.device)
            self._eval_stat = {}
        self._eval_x[:ng * self.mb].zero_()
        self._eval_x[:n].copy_(x)  # a short last batch is zero padded: eval-mode BN is per sample
        P = self.theta
        self.wprep[0](P)
        ops.stem_im2col(self._eval_x, self._eval_y, None, None, 0, 0, ng * self.mb, self.patches.hi, self.patches.lo,
                        self.labels_mb)

        def stat(v):
            rm, rv = self._bn_buffers(v.bn_name)
            # the same running statistics for every group
            mean = rm.unsqueeze(0).expand(ng, -1).contiguous()
            rstd = torch.rsqrt(rv + BN_EPS).unsqueeze(0).expand(ng, -1).contiguous()
            self._eval_stat[v.bn_name] = (mean, rstd)
            return mean, rstd, P.data_ptr() + 4 * v.gamma_off, P.data_ptr() + 4 * v.beta_off

        def bn_eval(u, out, second=None, res=None):
            rm, rs, ga, be = stat(u)
            sec = None
            if second is not None:
                rm2, rs2, ga2, be2 = stat(second)
                sec = (second.y, rm2, rs2, ga2, be2)
            ops.bn_apply(u.y, rm, rs, ga, be, u.Pg, u.cout, out.hi, out.lo, relu=True, second=sec, res=res, ng=ng)

        u = self.stem
        u.plan.forward(ng, 0, stats=False)
        bn_eval(u, u.out)
        for blk in self.blocks:
            last = len(blk.units) - 1
            for i, u in enumerate(blk.units):
                u.plan.forward(ng, 0, stats=False)
                if i < last:
                    bn_eval(u, u.out)
            u = blk.units[last]
            if blk.ds is not None:
                d = blk.ds
                if blk.pooled is not None:
                    xin = blk.x
                    ops.avgpool2_fwd(xin.hi, xin.lo, ng * self.mb, xin.h, xin.w, xin.c, blk.pooled.hi, blk.pooled.lo)
                d.plan.forward(ng, 0, stats=False)
                bn_eval(u, blk.out, second=d)
            else:
                bn_eval(u, blk.out, res=(blk.x.hi, blk.x.lo))
        a = self.last
        hi, lo = a.hi[:n], (a.lo[:n] if a.lo is not None else None)
        pooled = (hi.float() + (lo.float() if lo is not None else 0.0)).view(n, a.h * a.w, a.c).mean(dim=1)
        logits = torch.addmm(self._view(P, "fc.bias").view(-1), pooled,
                             self._view(P, "fc.weight").view(self.classes, a.c).t())
        return logits

    # ------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _num_passes(block_strength, mode, impl, acc):
        """train-mode forward passes of one group launch: pass 1 (unless the raw gradient is given) + the FD passes"""
        regularise = (block_strength != 0 or acc != 0) and mode != "raw"
        return (0 if mode == "reg" else 1) + ((2 if impl == "central" else 1) if regularise else 0)

    def _compute_ops(self, x_src, labels_src, perm, first, use_cursor, ng, block_strength, eps, mode="full",
                     impl="forward", acc=0.0, target="avg"):
        """The lane-local part of one group launch of ng microbatches: im2col, pass 1, gradient norms / eps_n, the
        finite-difference pass(es).  Touches only this lane's buffers (and disjoint entries of grad_norms), so launches
        of different lanes may run concurrently.
        mode "full": whole per-microbatch recipe; "raw": pass 1 only (training.py:76-83 + :162);
        "reg": regulariser only, self.g_n[0] already holds the raw gradient of the microbatch (modules.py:211-241).
        impl "forward" | "central" (modules.py:211-241 / :266-300); acc: acc_strength with root.pre_n as pre_grads;
        target "avg" | "pre" (the acc_strength pre-pass of training.py:128-142 accumulates raw gradients into pre_n)."""
        root = self.root
        mb, st, n = self.mb, self.stride, self.numel
        cur = self.cursor if use_cursor else None
        if x_src.dtype == torch.uint8:  # raw HWC dataset: crop / flip / normalise fused into the im2col
            ops.stem_im2col_u8aug(x_src, labels_src, perm, cur, first, mb, ng * mb, root.aug_params, root.aug_mean,
                                  root.aug_std, self.patches.hi, self.patches.lo, self.labels_mb)
        else:
            ops.stem_im2col(x_src, labels_src, perm, cur, first, mb, ng * mb, self.patches.hi, self.patches.lo,
                            self.labels_mb)
        passes = 0
        g, g2 = self.g_n, self.g2_n
        if mode != "reg":
            self._forward(ng, 0, self.theta, 0, g, S_LOSSG, S_CORRG, passes)
            self._backward(ng, 0, self.theta, 0, g)
            passes += 1
        norms = None
        if target == "avg":
            norms = root.grad_norms[root.norm_offset:] if root.norm_offset else root.grad_norms
        simple = acc == 0
        # |g_k|^2 -> grad_norms[k] (training.py:162) and, without acc_strength, eps_k = eps / (bs * |g_k|) (modules.py:223)
        ops.flat_sqnorm(g, n, self.sq_ws, self.scal, S_N2G, ng=ng, gstride=st, norms_out=norms, cursor=self.cursor,
                        eps_mode=1 if simple else 0, bs=block_strength, eps=eps, eps_base=S_EPSG)
        if (block_strength != 0 or acc != 0) and mode != "raw":
            pre = root.pre_n if acc != 0 else None
            if not simple:  # |bs*g + acc*pre|^2 -> eps_k (modules.py:217-223)
                ops.flat_sqnorm(g, n, self.sq_ws, self.scal, S_VSQG, ng=ng, gstride=st, y=pre, a=block_strength, b=acc,
                                eps_mode=2, eps=eps, eps_base=S_EPSG)
            points = ((1.0, g2),) if impl == "forward" else ((0.5, g2), (-0.5, self._g3()))
            for scale, gbuf in points:
                self.wprep[1](self.theta, ng=ng, grad=g, gstride=st, pre=pre, bs=block_strength, acc=acc, scale=scale,
                              scal=self.scal, eps_base=S_EPSG)
                ops.perturb_ranges(self.theta, g, st, pre, self.ranges, self.n_ranges, self.ranges_total,
                                   block_strength, acc, scale, self.scal, S_EPSG, self.theta_p, st, ng)
                self._forward(ng, 1, self.theta_p, st, gbuf, S_LOSS2G, S_CORR2G, passes)
                self._backward(ng, 1, self.theta_p, st, gbuf)
                passes += 1
        return passes

    def _commit_ops(self, ng, block_strength, accumulate, write_g, mode="full", impl="forward", acc=0.0, batch_clip=None,
                    target="avg", cursor_step=None):
        """The order-sensitive tail of a group launch: finite-difference combine + running mean of the ng microbatches in
        loader order (modules.py:232-240, training.py:45-47,166-168), the running-statistics EMA of every BatchNorm
        (pass 1 of microbatch k, pass 2 of k, pass 1 of k+1, ...) and the loss / accuracy sums.  Runs on the main stream,
        launch after launch, whichever lane computed the gradients."""
        root = self.root
        dst = root.avg_n if target == "avg" else root.pre_n
        st, n = self.stride, self.numel
        g, g2 = self.g_n, self.g2_n
        passes = self._num_passes(block_strength, mode, impl, acc)
        regularise = (block_strength != 0 or acc != 0) and mode != "raw"
        fused_mean = dst if (accumulate and batch_clip is None) else None
        if regularise:
            ops.fd_combine(g, g2, None if impl == "forward" else self._g3(), st, fused_mean, n, ng, self.scal, S_EPSG,
                           S_CF, self.cursor, write_g or batch_clip is not None)
            if accumulate and batch_clip is not None:
                ops.flat_sqnorm(g, n, self.sq_ws, self.scal, S_REGSQG, ng=ng, gstride=st)
                ops.mean_accumulate(g, st, dst, n, ng, self.cursor, self.scal, S_REGSQG, batch_clip, S_CLIPPED)
        elif accumulate:
            ops.mean_accumulate(g, st, dst, n, ng, self.cursor, self.scal, S_N2G, batch_clip or 0.0, S_CLIPPED)
        self.ema(passes, ng, BN_MOMENTUM)
        if mode != "reg":
            # the sums live in lane 0's scalars: commits run in loader order on the main stream, so the loss is added up
            # microbatch by microbatch whatever the group and lane counts are
            ops.group_finish(self.cursor, ng, self.scal, S_LOSS, S_CORRECT, S_LOSSG, S_CORRG, cursor_step=cursor_step,
                             totals=self.root.scal)

    def _group_ops(self, x_src, labels_src, perm, first, use_cursor, ng, block_strength, eps, accumulate, write_g,
                   mode="full", impl="forward", acc=0.0, batch_clip=None, target="avg"):
        """One whole group launch on the current stream (single-microbatch protocols)."""
        passes = self._compute_ops(x_src, labels_src, perm, first, use_cursor, ng, block_strength, eps, mode, impl, acc,
                                   target)
        self._commit_ops(ng, block_strength, accumulate, write_g, mode, impl, acc, batch_clip, target)
        return passes

    def _g3(self):
        if not hasattr(self, "g3_n"):
            self.g3_n = torch.zeros_like(self.g_n)
        return self.g3_n

    def _pre(self):
        root = self.root
        if not hasattr(root, "pre_n"):
            root.pre_n = torch.zeros(self.stride, device=self.device)[:self.numel]
            root.pre = torch.zeros(self.stride, device=self.device)[:self.numel]
        return root.pre_n

    def _graph_key(self, x_src, labels_src, perm, *rest):
        root = self.root
        return (x_src.data_ptr(), labels_src.data_ptr(), None if perm is None else perm.data_ptr(), *rest,
                root.grad_norms.data_ptr(), root.norm_offset,
                None if root.aug_params is None else root.aug_params.data_ptr(), tuple(root.aug_mean), tuple(root.aug_std))

    def _capture(self, parts, mode):
        """Warm-up run of the callables in `parts` (sets kernel attributes, loads modules) with the state restored
        afterwards, then ONE CUDA graph per part."""
        torch.cuda.synchronize()  # other lanes may still be writing shared entries
        state = self._save_state(mode)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for fn in parts:
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._restore_state(state)
        graphs = []
        for fn in parts:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                fn()
            graphs.append(graph)
        return graphs

    def _program(self, x_src, labels_src, perm, first, use_cursor, ng, block_strength, eps, accumulate=True,
                 write_g=False, use_graph=True, mode="full", impl="forward", acc=0.0, batch_clip=None, target="avg"):
        """Returns a callable running one whole group launch of ng microbatches on the current stream; captured into a
        CUDA graph on first use."""
        if acc != 0 or target == "pre":
            self._pre()
        if impl == "central":
            self._g3()
        args = (x_src, labels_src, perm, first, use_cursor, ng, block_strength, eps, accumulate, write_g, mode, impl, acc,
                batch_clip, target)
        if not use_graph:
            return lambda: self._group_ops(*args)
        key = self._graph_key(x_src, labels_src, perm, "whole", first, use_cursor, ng, float(block_strength), float(eps),
                              accumulate, write_g, mode, impl, float(acc), batch_clip, target)
        if key not in self._graphs:
            graphs = self._capture([lambda: self._group_ops(*args)], mode)
            self._graphs[key] = (graphs, (x_src, labels_src, perm))
        return self._graphs[key][0][0].replay

    def _lane_programs(self, x_src, labels_src, perm, first, use_cursor, ng, block_strength, eps, cursor_step,
                       use_graph=True, mode="full", impl="forward", acc=0.0, batch_clip=None, target="avg"):
        """(compute, commit) callables of one group launch on THIS lane: `compute` for the lane's stream, `commit` for the
        main stream; each captured into its own CUDA graph on first use."""
        if acc != 0 or target == "pre":
            self._pre()
        if impl == "central":
            self._g3()
        cargs = (x_src, labels_src, perm, first, use_cursor, ng, block_strength, eps, mode, impl, acc, target)
        margs = (ng, block_strength, True, False, mode, impl, acc, batch_clip, target, cursor_step)
        if not use_graph:
            return (lambda: self._compute_ops(*cargs)), (lambda: self._commit_ops(*margs))
        key = self._graph_key(x_src, labels_src, perm, "lane", first, use_cursor, ng, float(block_strength), float(eps),
                              cursor_step, mode, impl, float(acc), batch_clip, target)
        if key not in self._graphs:
            graphs = self._capture([lambda: self._compute_ops(*cargs), lambda: self._commit_ops(*margs)], mode)
            self._graphs[key] = (graphs, (x_src, labels_src, perm))
        compute, commit = self._graphs[key][0]
        return compute.replay, commit.replay

    def _save_state(self, mode):
        root = self.root
        bufs = [b.clone() for b in self.model.buffers()]
        return dict(avg=root.avg_n.clone(), scal=self.scal.clone(), totals=root.scal.clone(), cursor=self.cursor.clone(), bufs=bufs,
                    norms=root.grad_norms.clone(), g=self.g_n[0].clone() if mode == "reg" else None,
                    pre=root.pre_n.clone() if hasattr(root, "pre_n") else None)

    def _restore_state(self, st):
        root = self.root
        root.avg_n.copy_(st["avg"])
        root.scal.copy_(st["totals"])  # the loss / accuracy sums of every lane's commits live in lane 0's scalars
        self.scal.copy_(st["scal"])
        self.cursor.copy_(st["cursor"])
        root.grad_norms.copy_(st["norms"])
        if st["g"] is not None:
            self.g_n[0].copy_(st["g"])
        if st["pre"] is not None:
            root.pre_n.copy_(st["pre"])
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
            for lane in self.lanes:
                lane._graphs.clear()
        n = self.aug_params.shape[0]
        off = torch.randint(0, 2 * crop_padding + 1, (n, 2), device=self.device, generator=generator)
        flips = (torch.rand(n, device=self.device, generator=generator) < flip)
        self.aug_params[:, 0:2] = off.to(torch.int8)
        self.aug_params[:, 2] = flips.to(torch.int8)
        return self.aug_params

    def set_lr(self, lr):
        """correction factor lr/4 of modules.py:214, kept on the device so captured graphs stay valid"""
        for lane in self.lanes:
            lane.scal[S_CF] = lr / 4

    def begin_step(self, num_microbatches):
        if self.grad_norms is None or self.grad_norms.numel() < num_microbatches + self.G:
            self.grad_norms = torch.zeros(max(num_microbatches, 16) + self.G, device=self.device)
            for lane in self.lanes:
                lane._graphs.clear()
        self.grad_norms.zero_()
        self.avg_n.zero_()
        self._reset_sums()
        self.wprep[0](self.theta)  # theta is constant during the step: pass-1 operands once, not per launch

    def _reset_sums(self):
        for lane in self.lanes:
            lane.scal[S_LOSS:S_CORRECT + 1] = 0
            lane.scal[S_CLIPPED] = 0
            lane.cursor.zero_()

    def launch_sizes(self, count):
        return launch_sizes(count, self.G)

    def _begin_lanes(self):
        """A lane's cursor is the index (in loader order) of the first microbatch of its current launch; _launch sets it
        on the lane's stream once the lane's previous launch has been committed."""
        self._fork = torch.cuda.Event()
        self._fork.record(torch.cuda.current_stream())

    def _launch(self, j, ng, make, start, wait=None, after=None):
        """Group launch number j of a pass, microbatches start .. start + ng - 1: compute on lane j % lanes (its own
        stream, once its previous launch has been combined), commit on the current (main) stream -- commits therefore
        run in launch order."""
        lanes = self.lanes
        lane = lanes[j % len(lanes)]
        main = torch.cuda.current_stream()
        compute, commit = make(lane, ng, 0)
        if len(lanes) == 1:  # no concurrency: everything in stream order
            if wait is not None:
                main.wait_event(wait)
            lane.cursor.fill_(start)
            compute()
            if after is not None:
                after.record(main)
            commit()
            return
        with torch.cuda.stream(lane.stream):
            lane.stream.wait_event(self._fork)
            lane.stream.wait_event(lane.commit_done)
            if wait is not None:
                lane.stream.wait_event(wait)
            lane.cursor.fill_(start)
            compute()
            if after is not None:
                after.record(lane.stream)
            done = torch.cuda.Event()
            done.record(lane.stream)
        main.wait_event(done)
        commit()
        lane.commit_done.record(main)

    def _fold_lanes(self):
        """clip counters of the other lanes into lane 0's (device side, main stream; whole numbers: any order is exact).
        The loss / accuracy sums are accumulated in lane 0's scalars by every lane's commit (fb_group_finish totals)."""
        for lane in self.lanes[1:]:
            self.scal[S_CLIPPED] += lane.scal[S_CLIPPED]
            lane.scal[S_CLIPPED] = 0

    def _run_groups(self, count, make):
        """count microbatches as launches of launch_sizes(count) groups, alternating between the lanes;
        make(lane, ng, cursor_step) -> (compute, commit)"""
        self._begin_lanes()
        start = 0
        for j, ng in enumerate(self.launch_sizes(count)):
            self._launch(j, ng, make, start)
            start += ng
        self._fold_lanes()

    def accumulate_resident(self, X, Y, lr, block_strength, eps, first=0, count=None, perm=None, use_graph=True,
                            num_norms=None, norm_offset=0, implementation="forward-differences", acc_strength=0.0,
                            batch_clip=None, reduce_pre=None):
        """Full-batch accumulation over `count` consecutive microbatches of a device-resident dataset
        X [N,3,32,32] fp32 (or [N,32,32,3] uint8), Y [N] int64, starting at sample `first` (optionally through the index
        tensor `perm`).  Returns after enqueueing; results: self.avg (running mean, reference layout),
        self.grad_norms[:count], loss/correct sums in scal.  reduce_pre: callable applied to the mean raw gradient of
        the acc_strength pre-pass before it is used (the all-reduce of training.py:139-140)."""
        assert X.is_cuda and X.is_contiguous() and Y.dtype == torch.int64
        if X.dtype == torch.uint8:
            assert tuple(X.shape[1:]) == (32, 32, 3), "uint8 datasets are HWC [N,32,32,3]"
        else:
            assert X.dtype == torch.float32 and tuple(X.shape[1:]) == (3, 32, 32)
        n_avail = (perm.numel() if perm is not None else X.shape[0]) - first
        count = n_avail // self.mb if count is None else count
        if count < 0 or first < 0 or count * self.mb > n_avail:
            raise ValueError(f"{count} microbatches of {self.mb} from sample {first} exceed the {n_avail + first} samples "
                             "of the dataset (drop_last?)")
        impl = IMPLEMENTATIONS[implementation]
        self.begin_step(num_norms or count)
        self.norm_offset = int(norm_offset)
        self.set_lr(lr)
        if acc_strength != 0:
            # training.py:128-142: extra sweep for the mean raw gradient (pre_grads)
            self._pre().zero_()
            self._run_groups(count, lambda lane, ng, step: lane._lane_programs(
                X, Y, perm, first, True, ng, 0.0, eps, step, use_graph=use_graph, mode="raw", impl=impl,
                batch_clip=batch_clip, target="pre"))
            if reduce_pre is not None:
                reduce_pre(self.pre_n)
            self.from_native(self.pre_n, self.pre)
            self.bn_passes += count
            self._reset_sums()
        self._run_groups(count, lambda lane, ng, step: lane._lane_programs(
            X, Y, perm, first, True, ng, block_strength, eps, step, use_graph=use_graph, impl=impl, acc=acc_strength,
            batch_clip=batch_clip))
        self.from_native(self.avg_n, self.avg)
        regularised = block_strength != 0 or acc_strength != 0
        self.bn_passes += count * ((3 if impl == "central" else 2) if regularised else 1)
        return count

    def _stages(self):
        if not hasattr(self, "_x_stage"):
            n = self.G * self.mb
            k = len(self.lanes) + 1  # one staging buffer per lane in flight + one being filled
            self._x_stage = [torch.zeros(n, 3, 32, 32, device=self.device) for _ in range(k)]
            self._y_stage = [torch.zeros(n, device=self.device, dtype=torch.int64) for _ in range(k)]
            self._stage_free = [torch.cuda.Event() for _ in range(k)]
            self._copy_stream = torch.cuda.Stream()
        return self._x_stage, self._y_stage

    def accumulate_stream(self, loader, lr, block_strength, eps, num_microbatches, use_graph=True, norm_offset=0):
        """Full-batch accumulation over blocks (inputs [B,3,32,32], labels [B]) coming from a host-side iterable (the
        reference's DataLoader protocol, training.py:148-152): every block is split into microbatches
        (torch.chunk, training.py:155-156), copied host->device on a copy stream into one of the staging buffers of G
        microbatches and consumed by the captured graphs of a lane, so the copies of the next launches overlap the
        compute of the current ones."""
        xs, ys = self._stages()
        self.begin_step(num_microbatches)
        self.norm_offset = int(norm_offset)
        self.set_lr(lr)
        self._begin_lanes()
        mb, n_stage = self.mb, len(xs)
        state = dict(k=0, slot=0, launch=0, h2d=0, start=0)
        sizes = self.launch_sizes(num_microbatches)

        def flush():
            j, ng = state["launch"], state["slot"]
            if ng == 0:
                return
            i = j % n_stage
            ready = torch.cuda.Event()
            ready.record(self._copy_stream)
            self._launch(j, ng, lambda lane, g, step: lane._lane_programs(xs[i], ys[i], None, 0, False, g, block_strength,
                                                                          eps, step, use_graph=use_graph),
                         state["start"], wait=ready, after=self._stage_free[i])
            state["launch"], state["slot"], state["start"] = j + 1, 0, state["start"] + ng

        for inputs, labels in loader:
            chunks = max(labels.shape[0] // mb, 1)
            for xc, yc in zip(torch.chunk(inputs, chunks, dim=0), torch.chunk(labels, chunks, dim=0)):
                if xc.shape[0] != mb:
                    raise RuntimeError(f"microbatch of {xc.shape[0]} samples, engine built for {mb} (drop_last?)")
                i, j = state["launch"] % n_stage, state["slot"]
                with torch.cuda.stream(self._copy_stream):
                    if j == 0:
                        self._copy_stream.wait_event(self._stage_free[i])
                    xs[i][j * mb:(j + 1) * mb].copy_(xc, non_blocking=True)
                    ys[i][j * mb:(j + 1) * mb].copy_(yc, non_blocking=True)
                state["h2d"] += xc.numel() * xc.element_size() + yc.numel() * yc.element_size()
                state["k"] += 1
                state["slot"] = j + 1
                if state["slot"] == (sizes[state["launch"]] if state["launch"] < len(sizes) else self.G):
                    flush()
        flush()
        self._fold_lanes()
        self.from_native(self.avg_n, self.avg)
        self.bn_passes += state["k"] * (2 if block_strength != 0 else 1)
        self.h2d_bytes = state["h2d"]
        return state["k"]

    # ---- GradRegularizer / _compute_batched_gradient protocol ------------------------------------------------------
    def microbatch_gradient(self, inputs, labels):
        """training.py:76-83 on the device for one microbatch: fills self.g (raw gradient, reference layout), returns
        (loss, correct) as device scalars."""
        xs, ys = self._stages()
        xs[0][:self.mb].copy_(inputs)
        ys[0][:self.mb].copy_(labels)
        self.begin_step(1)
        self._program(xs[0], ys[0], None, 0, False, 1, 0.0, 0.0, accumulate=False, mode="raw")()
        self.from_native(self.g_n[0, :self.numel], self.g)
        self.bn_passes += 1
        return self.scal[S_LOSS], self.scal[S_CORRECT]

    def regularize(self, inputs, labels, lr, block_strength, eps, implementation="forward-differences",
                   acc_strength=0.0):
        """modules.py:211-300: self.g (raw gradient of this microbatch) <- regularised gradient, in place
        (acc_strength uses self.pre, see load_pre)."""
        xs, ys = self._stages()
        xs[0][:self.mb].copy_(inputs)
        ys[0][:self.mb].copy_(labels)
        if self.grad_norms is None:
            self.begin_step(1)
        self.cursor.zero_()
        self.set_lr(lr)
        impl = IMPLEMENTATIONS[implementation]
        self.to_native(self.g, self.g_n[0, :self.numel])
        self._program(xs[0], ys[0], None, 0, False, 1, block_strength, eps, accumulate=False, write_g=True, mode="reg",
                      impl=impl, acc=acc_strength)()
        self.from_native(self.g_n[0, :self.numel], self.g)
        self.bn_passes += 2 if impl == "central" else 1

    def load_pre(self, pre_grads):
        """pre_grads (list shaped like model.parameters(), training.py:128-142) -> flat self.pre (+ native copy)"""
        self._pre()
        for name, t in zip(self.names, pre_grads):
            self._view(self.pre, name).copy_(t.reshape(-1))
        self.to_native(self.pre, self.pre_n)

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
    def all_reduce_flat(self, flat, local_count, global_count):
        """fb_allreduce_flat of SURVEY.md 8b: flat <- sum over ranks of flat * local_count / global_count (the weights
        make the sum the exact global mean of per-rank running means), one NCCL all-reduce over NVLink."""
        import torch.distributed as dist

        ops.flat_scale(flat, flat.numel(), float(local_count) / float(global_count))
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)

    def all_reduce_mean(self, local_count, global_count):
        """avg holds the running mean over this rank's `local_count` microbatches; after this call every rank holds the
        mean over all `global_count` microbatches (exactly the single-process result up to fp32 summation order).
        The reference's own multi-process weighting (training.py:168 with num_machines > 1) is not a mean and is
        deliberately not reproduced."""
        import torch.distributed as dist

        self.all_reduce_flat(self.avg, local_count, global_count)
        pack = torch.cat([self.scal[S_LOSS:S_CORRECT + 1], self.scal[S_CLIPPED:S_CLIPPED + 1], self.grad_norms])
        dist.all_reduce(pack, op=dist.ReduceOp.SUM)
        self.scal[S_LOSS:S_CORRECT + 1] = pack[:2]
        self.scal[S_CLIPPED] = pack[2]
        self.grad_norms.copy_(pack[3:])

    def results(self, count):
        """Host read of the step scalars (one synchronisation): mean loss, correct count, grad_norms."""
        pack = torch.cat([self.scal[:16], self.grad_norms[:count]]).tolist()
        return dict(loss=pack[S_LOSS] / max(count, 1), correct=pack[S_CORRECT], loss_sum=pack[S_LOSS],
                    grad_norms=torch.tensor(pack[16:], dtype=torch.float32), clipped_batches=int(pack[S_CLIPPED]),
                    scal=pack[:16])

    def sync_bn_counters(self):
        """num_batches_tracked += number of train-mode passes (2 per microbatch with the regulariser)."""
        if self.bn_passes:
            for m in self.model.modules():
                if isinstance(m, torch.nn.BatchNorm2d):
                    m.num_batches_tracked += self.bn_passes
            self.bn_passes = 0

    def grads_list(self, flat):
        """Views of a flat buffer (reference layout) shaped like model.parameters()."""
        return [self._view(flat, n).view(self.shapes[n]) for n in self.names]
