"""Host-side operator layer over the C ABI: builds TMA descriptors / tap tables for each convolution and wraps the
layer kernels.  Everything here only prepares arguments; all arithmetic happens in the CUDA library (lib.py).

Layouts: activations NHWC; a "split" activation is a pair of bf16 planes (hi, lo) with x ~= hi + lo (lo is None in
plain-bf16 mode); conv outputs / activation gradients fp32 NHWC; dY bf16 NHWC; GEMM weights bf16 hi/lo as
wf[cout][tap*cin + ci] and wd[cin][tap*cout + co]; conv weight GRADIENTS in the kernels' native [co][tap][ci] layout.

Microbatch groups: every launch serves `ng` consecutive microbatches of `mb` images (include/fullbatch_b200.h).  All
geometry decisions that change the order of a floating-point reduction (N tile, BatchNorm partial rows, wgrad split-K,
BatchNorm-backward chunks) are functions of ONE group's problem and of the constant POLICY_GROUPS, never of `ng`, so the
result of a microbatch does not depend on how many microbatches share its launch.
"""
import ctypes as C
import os

import torch

from . import lib as L

NUM_SMS = 148
# nominal number of groups per launch the geometry policies are tuned for (a constant: results must not depend on ng)
POLICY_GROUPS = int(os.environ.get("FB_POLICY_GROUPS", "8"))

# Optional per-launch timing for bench.py's roofline: set PROFILE = [] to collect
# (family, algorithmic work, unit, start event, end event, label) around every wrapped launch on the current stream.
PROFILE = None
LAUNCHES = {"count": 0}  # kernels launched through this module (bench.py's gpu_launches claim)
_KERNELS_PER_CALL = {"fb_weight_prep_multi": 1, "fb_conv_gemm": 1, "fb_conv_wgrad": 1, "fb_reduce_multi": 1,
                     "fb_weight_prep": 1, "fb_stem_im2col": 1, "fb_stem_im2col_u8aug": 1, "fb_bn_stats": 2,
                     "fb_bn_apply": 1, "fb_bn_bwd": 2, "fb_bn_ema_multi": 1, "fb_avgpool2_fwd": 1, "fb_avgpool2_bwd": 1,
                     "fb_head_fwd_bwd": 3, "fb_flat_sqnorm": 2, "fb_perturb_ranges": 1, "fb_fd_combine": 1,
                     "fb_mean_accumulate": 1, "fb_group_finish": 1, "fb_flat_scale": 1, "fb_flat_relayout": 1,
                     "fb_sgd_step": 2}


def _call(family, work, unit, name, *args, label=None):
    LAUNCHES["count"] += _KERNELS_PER_CALL[name]
    if PROFILE is None:
        L.call(name, *args)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.call(name, *args)
    e1.record()
    PROFILE.append((family, work, unit, e0, e1, label or name))


def device_table(ctypes_array, device):
    """A ctypes array of structs -> uint8 tensor on the device (tables read by the *_multi kernels)."""
    return torch.frombuffer(bytearray(bytes(ctypes_array)), dtype=torch.uint8).to(device)


def pixel_tile(h, w):
    """128-pixel TMA box (tile_w, tile_h, tile_n) over an [n, h, w] pixel grid; the box always spans full rows."""
    if w > 128 or 128 % w != 0:
        raise RuntimeError(f"unsupported feature-map width {w} (no fallback path)")
    tile_h = min(h, 128 // w)
    if h % tile_h != 0 or 128 % (w * tile_h) != 0:
        raise RuntimeError(f"unsupported feature-map size {h}x{w}")
    return w, tile_h, 128 // (w * tile_h)


def tiles_per_group(mb, h, tile):
    """128-pixel tiles of one microbatch group (the C side's tiles_per_group)."""
    tw, th, tn = tile
    if tn == 1:
        return mb * (h // th)
    if mb % tn != 0:
        raise RuntimeError(f"a microbatch of {mb} images cannot be tiled by boxes of {tn} images ({h}x{tw} maps): the "
                           f"microbatch size must be a multiple of {tn}")
    return mb // tn


class MapSet:
    """A host array of CUtensorMap blobs plus references that keep the mapped tensors alive."""

    def __init__(self, n):
        self.n = n
        self.buf = C.create_string_buffer(L.FB_TMAP_BYTES * n)
        self.keep = []

    def slot(self, i):
        return C.addressof(self.buf) + i * L.FB_TMAP_BYTES

    @property
    def addr(self):
        return C.addressof(self.buf)


def encode_act(ms, i, plane, n, h, w, c, tile, phase=None):
    """Map slot i <- 4-D view of an NHWC bf16 plane [n,h,w,c]; phase=(ph,pw) selects a stride-2 sub-grid."""
    assert plane.dtype == torch.bfloat16 and plane.is_contiguous()
    tw, th, tn = tile
    base = plane.data_ptr()
    if phase is None:
        dims = (c, w, h, n)
        strides = (c, w * c, h * w * c)
    else:
        ph, pw = phase
        base += (ph * w + pw) * c * 2
        dims = (c, w // 2, h // 2, n)
        strides = (2 * c, 2 * w * c, h * w * c)
    L.check(L.load().fb_tmap_encode_act4d(ms.slot(i), base, dims[0], dims[1], dims[2], dims[3], strides[0], strides[1],
                                          strides[2], 64, tw, th, tn), "fb_tmap_encode_act4d")
    ms.keep.append(plane)


def encode_mat(ms, i, mat, k, rows, box_rows):
    assert mat.dtype == torch.bfloat16 and mat.is_contiguous() and mat.dim() == 2
    L.check(L.load().fb_tmap_encode_mat2d(ms.slot(i), mat.data_ptr(), k, rows, mat.stride(0), 64, box_rows),
            "fb_tmap_encode_mat2d")
    ms.keep.append(mat)


def choose_n_tile(m_tiles, n_total, a_planes=2, b_planes=2):
    """N tile of the generic GEMM by a small cost model: waves of the persistent grid x tensor-pipe clocks per K=16
    step of one tile.  An SS-mode MMA costs max(64, N/2) clocks (the 128x16 A tile is re-read from shared memory per
    instruction), a 64-wide tile with split weights issues `a_planes` stacked N=128 instructions, otherwise one
    instruction per operand-plane combination."""
    combos = 3 if (a_planes == 2 and b_planes == 2) else a_planes * b_planes
    best, best_cost = 64, None
    for nt in (64, 128, 256):
        if n_total % nt != 0:
            continue
        if nt == 64 and b_planes == 2:
            clocks = a_planes * 64
        else:
            clocks = combos * max(64, nt // 2)
        tiles = m_tiles * (n_total // nt)
        waves = -(-tiles // NUM_SMS)
        cost = waves * clocks
        if best_cost is None or cost < best_cost or (cost == best_cost and nt > best):
            best, best_cost = nt, cost
    return best


def _s2_tap(k):
    """3x3 / stride 2 / pad 1: input index 2*o + k - 1 -> (phase, shift in the phase grid)."""
    return ((1, -1), (0, 0), (1, 0))[k]


class ConvGemm:
    """One fb_conv_gemm problem with frozen descriptors; __call__(ng) launches it for the first ng groups."""

    def __init__(self, a_maps, b_maps, n_phases, a_planes, b_planes, taps, cblocks, tile, grid_h, mb, n_total, out,
                 out_strides, accumulate, n_tile, b_group_rows=0, tapgroups=None, reverse=False, cta_pair=False,
                 halo=False, policy_groups=1, sched_k_iters=0):
        """tapgroups: optional [(tap0, n_taps, out_off_elements)] -- tap groups that run over the same pixel grid and
        write to different offsets (the four output phases of a stride-2 dgrad in one launch)."""
        self.a_maps, self.b_maps = a_maps, b_maps
        args = L.ConvGemmArgs()
        args.host_a_maps, args.host_b_maps = a_maps.addr, b_maps.addr
        args.n_phases, args.a_planes, args.b_planes = n_phases, a_planes, b_planes
        if len(taps) > L.FB_MAX_TAPS:
            raise RuntimeError("too many taps")
        args.n_taps, args.cblocks = len(taps), cblocks
        for i, (phase, dh, dw, k0) in enumerate(taps):
            args.taps[i] = L.Tap(phase, dh, dw, 0, k0)
        args.tile_w, args.tile_h, args.tile_n = tile
        args.grid_h = grid_h
        args.n_total, args.n_tile = n_total, n_tile
        self.out = out
        args.out = out.data_ptr()
        args.out_sn, args.out_sh, args.out_sw = out_strides
        args.accumulate = int(accumulate)
        if tapgroups:
            if len(tapgroups) > 4:
                raise RuntimeError("at most 4 tap groups")
            args.n_tapgroups = len(tapgroups)
            for i, (tap0, n_taps, off) in enumerate(tapgroups):
                args.tapgroups[i] = L.TapGroup(tap0, n_taps, off)
        args.mg_imgs = mb
        args.cta_pair = int(cta_pair)
        args.halo = int(halo)
        args.policy_groups = int(policy_groups)
        args.sched_k_iters = int(sched_k_iters)
        args.b_group_rows = b_group_rows
        args.reverse = int(reverse)
        self.mb = mb
        self.args = args
        self.flops_per_group = 0.0  # algorithmic FLOPs of one group (set by Conv2dPlan)
        self.label = "conv_gemm"
        self.stats = None

    def set_stats(self, stats_ws, tickets, mean, rstd, eps):
        """BatchNorm statistics of the output in the epilogue (per group): see fb_conv_gemm_args."""
        self.stats = (stats_ws, tickets, mean, rstd)
        self._stats_ptrs = (stats_ws.data_ptr(), tickets.data_ptr(), mean.data_ptr(), rstd.data_ptr())
        self.args.bn_eps = eps

    def __call__(self, ng=1, bn_batch=None, stats=True, reverse=False):
        a = self.args
        a.ng, a.grid_n = ng, ng * self.mb
        a.reverse = int(reverse)
        if self.stats is not None and stats:
            a.stats_ws, a.tickets, a.bn_mean, a.bn_rstd = self._stats_ptrs
            a.bn_batch = bn_batch
        else:
            a.stats_ws = None
        _call("conv_gemm", self.flops_per_group * ng, "flop", "fb_conv_gemm", C.byref(a), label=self.label)


class Conv2dPlan:
    """Forward / dgrad / wgrad launches of one bias-free Conv2d (k in {1,3}, stride in {1,2}, pad (k-1)/2) on NHWC data
    holding up to G groups of mb images.

    x_hi/x_lo : [G*mb,h,w,cin] bf16 planes (conv input; for the stem: the im2col patches with cin=64, k=1)
    y         : [G*mb,ho,wo,cout] fp32 (forward output)
    dy        : [G*mb,ho,wo,cout] bf16 (output gradient)
    dx        : [G*mb,h,w,cin] fp32 (input gradient) or None if no dgrad is needed
    wsets     : [(wf_hi, wf_lo, wd_hi, wd_lo)] x 2: [0] shared weights ([cout][K] / [cin][K']), [1] per-group weights
                ([G*cout][K] / [G*cin][K']: the perturbed points of the finite-difference passes)
    w_offset  : element offset of the weight (gradient) inside the flat buffers; grad_ld: row length of the flat
                gradient (taps*cin; 27 for the stem, whose GEMM K is padded to 64)
    """

    def __init__(self, mb, G, h, w, cin, cout, k, stride, x_hi, x_lo, y, dy, dx, wsets, w_offset, split=True, alg_k=None,
                 grad_cols=None, bn=None, allow_pair=True, policy_groups=None, fwd_hi_only=False):
        """fwd_hi_only: the forward GEMM reads only the hi plane of the activations (2 products per MAC instead of 3:
        drops x_lo * w_hi; numerics ablation, engine precision "split_w")."""
        assert k in (1, 3) and stride in (1, 2) and cin % 64 == 0 and cout % 64 == 0
        assert not (k == 1 and stride == 2)
        self.mb, self.G, self.h, self.w, self.cin, self.cout, self.k, self.stride = mb, G, h, w, cin, cout, k, stride
        # number of groups per launch the geometry is tuned for (never the ng of a launch: results must not depend on it)
        policy = int(policy_groups or POLICY_GROUPS)
        n = G * mb
        ho, wo = h // stride, w // stride
        self.ho, self.wo = ho, wo
        taps = k * k
        self.taps = taps
        self.w_offset = w_offset
        # algorithmic FLOPs of one forward (= one dgrad = one wgrad) of ONE group: 2 * pixels * cout * K
        self.alg_flops = 2.0 * mb * ho * wo * cout * (alg_k or taps * cin)
        x_lo = x_lo if split else None
        planes = 2 if x_lo is not None else 1
        wplanes = 2 if (split and wsets[0][1] is not None) else 1
        if planes != wplanes:
            raise RuntimeError("activation and weight operands must both be split or both be plain bf16")
        tile = pixel_tile(ho, wo)
        self.tile = tile
        cb_in, cb_out = cin // 64, cout // 64
        mtg = tiles_per_group(mb, ho, tile)
        self.mtg = mtg

        # ---- activation maps over the conv input, indexed [phase * planes + plane]
        nph = 4 if stride == 2 else 1
        xs = MapSet(nph * planes)
        for p in range(nph):
            for pl, t in enumerate((x_hi, x_lo)[:planes]):
                encode_act(xs, p * planes + pl, t, n, h, w, cin, tile, phase=(p // 2, p % 2) if stride == 2 else None)
        self.x_maps = xs

        def tap_geom(kh, kw):
            if k == 1:
                return 0, 0, 0
            if stride == 1:
                return 0, kh - 1, kw - 1
            (ph, dh), (pw, dw) = _s2_tap(kh), _s2_tap(kw)
            return ph * 2 + pw, dh, dw

        # ---- forward (one ConvGemm per weight set)
        fplanes = 1 if fwd_hi_only else planes  # operand planes of the forward A operand
        xs_fwd = xs
        if fplanes != planes:
            xs_fwd = MapSet(nph)
            for p in range(nph):
                encode_act(xs_fwd, p, x_hi, n, h, w, cin, tile, phase=(p // 2, p % 2) if stride == 2 else None)
            self.x_maps_fwd = xs_fwd
        n_tile = choose_n_tile(mtg * policy, cout, fplanes, wplanes)
        self.n_tile = n_tile
        ftaps = []
        for kh in range(k):
            for kw in range(k):
                phase, dh, dw = tap_geom(kh, kw)
                ftaps.append((phase, dh, dw, (kh * k + kw) * cin))
        # haloed A boxes: the three taps of a filter column share one box of tile_h + 2 rows (taps dw-major)
        halo = self._use_halo(allow_pair, k, stride, tile, n_tile, fplanes)
        self.halo_fwd = halo
        if halo:
            ftaps = [(0, kh - 1, kw - 1, (kh * k + kw) * cin) for kw in range(k) for kh in range(k)]
            xs_f = MapSet(fplanes)
            for pl, t in enumerate((x_hi, x_lo)[:fplanes]):
                encode_act(xs_f, pl, t, n, h, w, cin, (tile[0], tile[1] + 2, tile[2]))
        else:
            xs_f = xs_fwd
        self.fwd = []
        # CTA pairs (M = 256 tiles over two SMs): each CTA fetches half of every weight tile -> half-height B boxes
        k_iters = taps * cb_in if bn is not None else 0  # tile schedule of launches that collect statistics
        pair = 0 if halo else self._use_pair(allow_pair, mtg, cout, n_tile, fplanes, policy, k_iters)
        self.pair_fwd = pair
        for si, (wf_hi, wf_lo, _, _) in enumerate(wsets):
            bs = MapSet(wplanes)
            rows = cout * (G if si == 1 else 1)
            for pl, t in enumerate((wf_hi, wf_lo)[:wplanes]):
                encode_mat(bs, pl, t, taps * cin, rows, n_tile // 2 if pair else n_tile)
            g = ConvGemm(xs_f, bs, nph, fplanes, wplanes, ftaps, cb_in, tile, ho, mb, cout, y,
                         (ho * wo * cout, wo * cout, cout), False, n_tile, b_group_rows=cout if si == 1 else 0,
                         cta_pair=pair, halo=halo, policy_groups=policy, sched_k_iters=k_iters)
            g.flops_per_group = self.alg_flops
            g.label = f"fwd{si} {h}x{w} {cin}->{cout} k{k}s{stride} nt{n_tile}{('', ' pair', ' mcast')[pair]}{' halo' if halo else ''}"
            self.fwd.append(g)
        # BatchNorm statistics fused into the forward epilogue
        self.stat_rows = L.load().fb_conv_stats_rows(mtg, cout // n_tile, policy, k_iters)
        if bn is not None:
            mean, rstd, eps = bn
            dev = y.device
            self.stats_ws = torch.zeros(G, self.stat_rows, 2, cout, device=dev)
            self.tickets = torch.zeros(G * (cout // n_tile), device=dev, dtype=torch.int32)
            for g in self.fwd:
                g.set_stats(self.stats_ws, self.tickets, mean, rstd, eps)

        # ---- dgrad
        self.dgrads = []
        if dx is not None:
            dys = MapSet(1)
            encode_act(dys, 0, dy, n, ho, wo, cout, tile)
            self.dy_maps_d = dys
            m_tiles_d = mtg * (4 if stride == 2 else 1)  # the four output phases of a stride-2 dgrad share one launch
            n_tile_d = choose_n_tile(m_tiles_d * policy, cin, 1, wplanes)
            if stride == 1:
                dtaps, tapgroups = [], None
                for kh in range(k):
                    for kw in range(k):
                        dh, dw = (1 - kh, 1 - kw) if k == 3 else (0, 0)
                        dtaps.append((0, dh, dw, (kh * k + kw) * cout))
                strides = (h * w * cin, w * cin, cin)
            else:
                # stride 2: output pixel (2i+ph, 2j+pw) gathers taps kh with (ph + 1 - kh) even: ho = i + (ph+1-kh)/2
                def taps_for(par):
                    return [(1, 0)] if par == 0 else [(0, 1), (2, 0)]  # (k index, shift in the dY grid)

                # one launch: four tap groups = the four output phases (1 + 2 + 2 + 4 taps)
                dtaps, tapgroups = [], []
                for ph in range(2):
                    for pw in range(2):
                        tap0 = len(dtaps)
                        for kh, dh in taps_for(ph):
                            for kw, dw in taps_for(pw):
                                dtaps.append((0, dh, dw, (kh * 3 + kw) * cout))
                        tapgroups.append((tap0, len(dtaps) - tap0, (ph * w + pw) * cin))
                strides = (h * w * cin, 2 * w * cin, 2 * cin)
            halo_d = self._use_halo(allow_pair, k, stride, tile, n_tile_d, 1)
            self.halo_dgrad = halo_d
            if halo_d:
                dtaps = [(0, 1 - kh, 1 - kw, (kh * k + kw) * cout) for kw in (2, 1, 0) for kh in (2, 1, 0)]
                dys_d = MapSet(1)
                encode_act(dys_d, 0, dy, n, ho, wo, cout, (tile[0], tile[1] + 2, tile[2]))
            else:
                dys_d = dys
            pair_d = 0 if halo_d else self._use_pair(allow_pair, mtg, cin, n_tile_d, 1, policy, 0)
            self.pair_dgrad = pair_d
            for si, (_, _, wd_hi, wd_lo) in enumerate(wsets):
                ds = MapSet(wplanes)
                rows = cin * (G if si == 1 else 1)
                for pl, t in enumerate((wd_hi, wd_lo)[:wplanes]):
                    encode_mat(ds, pl, t, taps * cout, rows, n_tile_d // 2 if pair_d else n_tile_d)
                g = ConvGemm(dys_d, ds, 1, 1, wplanes, dtaps, cb_out, tile, ho, mb, cin, dx, strides, False, n_tile_d,
                             b_group_rows=cin if si == 1 else 0, tapgroups=tapgroups, cta_pair=pair_d, halo=halo_d,
                             policy_groups=policy)
                g.flops_per_group = self.alg_flops
                g.label = f"dgrad{si} {h}x{w} {cin}->{cout} k{k}s{stride} nt{n_tile_d}{('', ' pair', ' mcast')[pair_d]}{' halo' if halo_d else ''}"
                self.dgrads.append(g)

        # ---- wgrad
        wa = L.WgradArgs()
        dym = MapSet(1)
        encode_act(dym, 0, dy, n, ho, wo, cout, tile)
        self.dy_map_w = dym
        n_slots = taps * cb_in
        co_tiles = -(-cout // 128)
        halo_w = self._wgrad_halo(k, stride, tile, planes)
        if halo_w:
            # taps in triples that share dw: per pixel block ONE haloed X box (tile_h + 2 rows) serves dh = -1, 0, +1
            hx = MapSet(planes)
            for pl, t in enumerate((x_hi, x_lo)[:planes]):
                encode_act(hx, pl, t, n, h, w, cin, (tile[0], tile[1] + 2, 1))
            self.x_maps_halo = hx
            wa.host_x_maps, wa.n_x_maps = hx.addr, hx.n
            i = 0
            for dwi in (-1, 0, 1):
                for dhi in (-1, 0, 1):
                    wa.taps[i] = L.WgradTap(0, dhi, dwi, (dhi + 1) * 3 + (dwi + 1))
                    i += 1
            spc = 3
            wa.halo = 1
        else:
            wa.host_x_maps, wa.n_x_maps = xs.addr, xs.n
            for kh in range(k):
                for kw in range(k):
                    phase, dh, dw = tap_geom(kh, kw)
                    wa.taps[kh * k + kw] = L.WgradTap(phase, dh, dw, kh * k + kw)
            spc = self._slots_per_cta(n_slots, planes)
        wa.host_dy_map = dym.addr
        wa.planes = planes
        wa.n_taps, wa.cblocks = taps, cb_in
        self.splits = self.wgrad_splits(mtg, co_tiles, -(-n_slots // spc), policy)
        wa.slots_per_cta = spc
        wa.cout, wa.cin = cout, cin
        wa.tile_w, wa.tile_h, wa.tile_n = tile
        wa.grid_h = ho
        wa.splits = self.splits
        wa.mg_imgs = mb
        self.k_ld = taps * cin                      # row length of the kernel's output matrix
        self.grad_cols = grad_cols or self.k_ld     # row length of the flat gradient (27 for the stem)
        # without split-K (and with matching rows) the epilogue writes the flat gradient itself
        self.direct = self.splits == 1 and self.grad_cols == self.k_ld
        self.partial = None
        self.wargs = wa

    @staticmethod
    def _use_halo(allow_pair, k, stride, tile, n_tile, a_planes):
        """Haloed A boxes (fb_conv_gemm_args.halo) for 3x3 / stride-1 convolutions whose 128-pixel tiles are whole rows
        of one image (32x32: 4 rows, 16x16: 8 rows) and whose N tile leaves room for the two boxes (<= 128).
        B200 measurements (profiles/r2_conv_limits.txt): the conv GEMMs run within 5-20 % of the time they take with
        ALL operand loads switched off (the tcgen05 instruction stream is the bound, not L2), so fetching a third to a
        half fewer operand bytes is worth only ~1 % of the step (39.0k -> 39.45k images/s in a 50k-image run, mostly
        the 64-wide split forward: 257 -> 243 us).  allow_pair = "halo" / "nohalo" force it on / off (tests);
        FB_HALO=0 disables."""
        ok = k == 3 and stride == 1 and tile[2] == 1 and n_tile <= 128
        if not ok or allow_pair == "nohalo" or allow_pair == "force" or allow_pair == "mcast" or not allow_pair:
            return False
        return allow_pair == "halo" or os.environ.get("FB_HALO", "1") != "0"

    @staticmethod
    def _use_pair(allow_pair, mtg, n_total, n_tile, a_planes, policy_groups, sched_k_iters):
        """Cluster mode of a conv GEMM (fb_conv_gemm_args.cta_pair): 0 = independent CTAs, 1 = CTA pairs (cta_group::2),
        2 = clusters of two CTAs that share every weight tile by TMA multicast (bit-identical to 0).
        A pair cannot stack the hi / lo weight planes into one wide instruction, so it only wins where nothing is
        stacked anyway (256-wide tiles) or where the stacked instruction is replaced one for one (single-plane A
        operand = dgrad, 128-wide tiles).  Measured on B200 (profiles/r2_cta_pairs.txt): 4x4x512 forward 215 -> 166 us,
        16x16x128 dgrad 117 -> 106 us, but 32x32x64 forward 258 -> 289 us.  The multicast mode halves the weight bytes a
        CTA pulls from L2, which turned out not to be the bound (FB_MCAST=1 enables it where the schedule allows).  allow_pair = "force" / "mcast": that mode wherever
        the schedule allows (tests)."""
        if not allow_pair or not L.load().fb_conv_pair_ok(mtg, n_total // n_tile, policy_groups, sched_k_iters):
            return 0
        if allow_pair == "mcast":
            return 2
        if allow_pair == "force" or os.environ.get("FB_PAIR_ALL") == "1":
            return 1
        if n_tile == 256 or (a_planes == 1 and n_tile == 128):
            return 1
        return 2 if os.environ.get("FB_MCAST", "0") == "1" else 0  # measured within +-2 % of independent CTAs: off

    @staticmethod
    def wgrad_splits(pixel_blocks_per_group, co_tiles, slot_groups, policy_groups=None):
        """split-K of ONE group's pixel blocks so that `policy_groups` groups fill one wave of the 148 SMs"""
        ctas = co_tiles * slot_groups * int(policy_groups or POLICY_GROUPS)
        return max(1, min(pixel_blocks_per_group, NUM_SMS // ctas))

    @staticmethod
    def _wgrad_halo(k, stride, tile, planes):
        """Haloed X boxes for wgrad: 3x3 / stride 1 on maps whose 128-pixel tiles are whole rows of one image with
        1024-byte-aligned rows, two stages of `planes` boxes within the 128 KB X ring (fb_wgrad_args.halo)."""
        if os.environ.get("FB_WGRAD_HALO", "1") != "1" or k != 3 or stride != 1 or tile[2] != 1:
            return False
        return (tile[0] * 128) % 1024 == 0 and planes * (tile[1] + 2) * tile[0] * 128 * 2 <= 128 * 1024

    @staticmethod
    def _slots_per_cta(n_slots, planes):
        """(tap, ci-block) accumulators per CTA: each takes 64*planes of the 512 TMEM columns."""
        cap = 8 // planes
        if n_slots <= cap:
            return n_slots
        return 3 if n_slots == 9 else cap

    def partial_elems(self):
        """fp32 elements of split-K workspace this layer needs (0: the epilogue writes the gradient directly)."""
        return 0 if self.direct else self.G * self.splits * self.cout * self.k_ld

    def bind_partial(self, partial):
        """partial: this layer's slice of the workspace -> the entry of the reduce table (src, strides, shape)"""
        assert not self.direct and partial.numel() >= self.partial_elems()
        self.partial = partial
        vec = 4 if (self.grad_cols % 4 == 0 and self.k_ld % 4 == 0 and self.w_offset % 4 == 0) else 1
        n_blocks = -(-(self.cout * self.grad_cols) // (256 * vec))
        return dict(src=partial.data_ptr(), src_gstride=self.splits * self.cout * self.k_ld,
                    src_sstride=self.cout * self.k_ld, dst_off=self.w_offset, dst_ld=self.grad_cols, splits=self.splits,
                    rows=self.cout, cols=self.grad_cols, src_ld=self.k_ld, n_blocks=n_blocks, vec=vec)

    def forward(self, ng, wset, bn_batch=None, stats=True):
        self.fwd[wset](ng, bn_batch, stats)

    def dgrad(self, ng, wset, reverse=False):
        self.dgrads[wset](ng, reverse=reverse)

    def wgrad(self, ng, gbuf, gstride):
        """weight gradient of the first ng groups -> gbuf[g*gstride + w_offset ...] (native layout) or the split-K
        workspace (then fb_reduce_multi finishes all layers at once)"""
        wa = self.wargs
        wa.ng, wa.grid_n = ng, ng * self.mb
        if self.direct:
            wa.out = gbuf.data_ptr() + self.w_offset * 4
            wa.out_gstride, wa.out_sstride = gstride, 0
        else:
            wa.out = self.partial.data_ptr()
            wa.out_gstride, wa.out_sstride = self.splits * self.cout * self.k_ld, self.cout * self.k_ld
        _call("conv_wgrad", self.alg_flops * ng, "flop", "fb_conv_wgrad", C.byref(wa),
              label=f"wgrad {self.h}x{self.w} {self.cin}->{self.cout} k{self.k}s{self.stride} splits{self.splits}"
                    f"{' halo' if wa.halo else ''}")


class ReduceTable:
    """Device table of fb_reduce_multi: the split-K reductions of all layers in one launch."""

    def __init__(self, entries, device):
        arr = (L.ReduceEntry * len(entries))()
        block, nbytes = 0, 0.0
        for i, e in enumerate(entries):
            arr[i] = L.ReduceEntry(e["src"], e["src_gstride"], e["src_sstride"], e["dst_off"], e["dst_ld"], e["splits"],
                                   e["rows"], e["cols"], e["src_ld"], block, e["n_blocks"], e["vec"], 0)
            block += e["n_blocks"]
            nbytes += 4.0 * e["rows"] * e["cols"] * (e["splits"] + 1)
        self.n, self.blocks, self.bytes_per_group = len(entries), block, nbytes
        self.table = device_table(arr, device)

    def __call__(self, ng, gbuf, gstride):
        _call("wgrad_reduce", self.bytes_per_group * ng, "byte", "fb_reduce_multi", self.table.data_ptr(), self.n,
              self.blocks, gbuf.data_ptr(), gstride, ng)


class WeightPrepTable:
    """Device table for fb_weight_prep_multi: entries = (w_offset, cout, cin, taps, wf_hi, wf_lo, wd_hi, wd_lo);
    per_group: the operand matrices hold G row blocks (one per microbatch group)."""

    def __init__(self, entries, device, per_group):
        arr = (L.WprepEntry * len(entries))()
        block, nbytes = 0, 0.0
        for i, (off, cout, cin, taps, wf_hi, wf_lo, wd_hi, wd_lo) in enumerate(entries):
            nb = (cout // 32) * (cin // 32) if cin % 32 == 0 else 4
            kf = wf_hi.stride(0)
            kd = wd_hi.stride(0) if wd_hi is not None else 0
            arr[i] = L.WprepEntry(off, cout, cin, taps, block, nb, 0, wf_hi.data_ptr(), L.ptr(wf_lo), L.ptr(wd_hi),
                                  L.ptr(wd_lo), kf, kd, cout * kf if per_group else 0, cin * kd if per_group else 0)
            block += nb
            planes = (1 + (wf_lo is not None)) * (1 + (wd_hi is not None))
            nbytes += cout * cin * taps * (4.0 + (4.0 if per_group else 0.0) + 2.0 * planes)
        self.keep = entries
        self.n, self.blocks, self.bytes_per_group = len(entries), block, nbytes
        self.table = device_table(arr, device)

    def __call__(self, theta, ng=1, grad=None, gstride=0, pre=None, bs=0.0, acc=0.0, scale=1.0, scal=None, eps_base=0):
        _call("weight_prep", self.bytes_per_group * ng, "byte", "fb_weight_prep_multi", theta.data_ptr(),
              self.table.data_ptr(), self.n, self.blocks, L.ptr(grad), gstride, L.ptr(pre), bs, acc, scale, L.ptr(scal),
              eps_base, ng)


class BnEmaTable:
    """Device table for fb_bn_ema_multi: (running_mean, running_var, batch [passes][G][2][C]) per BatchNorm layer."""

    def __init__(self, entries, device):
        arr = (L.BnEmaEntry * len(entries))()
        c0 = 0
        for i, (rm, rv, batch, pass_stride, Cc) in enumerate(entries):
            arr[i] = L.BnEmaEntry(rm.data_ptr(), rv.data_ptr(), batch.data_ptr(), pass_stride, Cc, c0)
            c0 += Cc
        self.keep = entries
        self.n, self.channels = len(entries), c0
        self.table = device_table(arr, device)

    def __call__(self, n_passes, ng, momentum):
        _call("misc", 4.0 * self.channels * (4 + 2 * n_passes * ng), "byte", "fb_bn_ema_multi", self.table.data_ptr(),
              self.n, self.channels, n_passes, ng, momentum)


# ---- thin wrappers of the layer kernels ------------------------------------------------------------------------------

def weight_prep(w_oihw, cout, cin, taps, wf_hi, wf_lo, wd_hi=None, wd_lo=None):
    planes = (1 + (wf_lo is not None)) * (1 + (wd_hi is not None))
    _call("weight_prep", cout * cin * taps * (4.0 + 2.0 * planes), "byte", "fb_weight_prep", w_oihw.data_ptr(), cout, cin,
          taps, wf_hi.data_ptr(), L.ptr(wf_lo), wf_hi.stride(0),
          L.ptr(wd_hi), L.ptr(wd_lo), wd_hi.stride(0) if wd_hi is not None else 0)


def stem_im2col(x, labels, perm, cursor, first, cursor_stride, n, p_hi, p_lo, labels_out):
    _call("stem_im2col", n * (3072 * 4.0 + 1024 * 64 * 2.0 * (1 + (p_lo is not None))), "byte", "fb_stem_im2col",
          x.data_ptr(), L.ptr(labels), L.ptr(perm), L.ptr(cursor), first, cursor_stride, n, p_hi.data_ptr(),
          L.ptr(p_lo), L.ptr(labels_out))


def stem_im2col_u8aug(x_u8, labels, perm, cursor, first, cursor_stride, n, aug, mean, std, p_hi, p_lo, labels_out):
    """x_u8: [N,32,32,3] uint8 on the device; aug: int8 [.,4] (dx, dy, flip, 0) per position or None; mean/std: 3 floats"""
    m = (L.f32 * 3)(*[float(v) for v in mean])
    sd = (L.f32 * 3)(*[float(v) for v in std])
    _call("stem_im2col", n * (3072.0 + 1024 * 64 * 2.0 * (1 + (p_lo is not None))), "byte", "fb_stem_im2col_u8aug",
          x_u8.data_ptr(), L.ptr(labels), L.ptr(perm), L.ptr(cursor), first, cursor_stride, n, L.ptr(aug), m, sd,
          p_hi.data_ptr(), L.ptr(p_lo), L.ptr(labels_out))


def bn_stats(y, P, Cc, ws, mean, rstd, running_mean, running_var, momentum=0.1, eps=1e-5):
    _call("bn_stats", 4.0 * P * Cc, "byte", "fb_bn_stats", y.data_ptr(), P, Cc, ws.data_ptr(), mean.data_ptr(),
          rstd.data_ptr(), L.ptr(running_mean), L.ptr(running_var), momentum, eps)


def bn_apply(y, mean, rstd, gamma, beta, P, Cc, out_hi, out_lo, relu=True, second=None, res=None, ng=1, param_gstride=0,
             reverse=False, mask_out=None):
    """P: pixels per group; mean / rstd: [ng][C]; gamma / beta: pointers (ints) or tensors, + g*param_gstride per group;
    second = (y2, mean2, rstd2, gamma2, beta2)"""
    def p(t):
        return t if isinstance(t, int) or t is None else t.data_ptr()

    a = L.BnApplyArgs()
    a.y, a.mean, a.rstd, a.gamma, a.beta = (p(t) for t in (y, mean, rstd, gamma, beta))
    if second is not None:
        a.y2, a.mean2, a.rstd2, a.gamma2, a.beta2 = (p(t) for t in second)
    if res is not None:
        a.res_hi, a.res_lo = res[0].data_ptr(), L.ptr(res[1])
    a.relu, a.P, a.C = int(relu), P, Cc
    a.out_hi, a.out_lo = out_hi.data_ptr(), L.ptr(out_lo)
    a.ng, a.param_gstride, a.reverse = ng, param_gstride, int(reverse)
    a.mask_out = L.ptr(mask_out)
    planes = 1 + (out_lo is not None)
    per_elem = 4.0 + (4.0 if second is not None else 0.0) + (2.0 * planes if res is not None else 0.0) + 2.0 * planes + \
        (0.125 if mask_out is not None else 0.0)
    _call("bn_fwd", per_elem * P * Cc * ng, "byte", "fb_bn_apply", C.byref(a),
          label=f"bn_apply P{P} C{Cc} {per_elem:.0f}B/elem")


def bn_bwd_ws_floats(P, Cc, G, policy_groups=None):
    chunks = L.load().fb_bn_bwd_chunks(P, Cc, int(policy_groups or POLICY_GROUPS))
    if chunks <= 0:
        raise RuntimeError(f"fb_bn_bwd does not support {Cc} channels")
    return 16 + G * (2 * Cc * chunks + 2 * Cc)


def bn_bwd(dA, mask_hi, y, mean, rstd, gamma, P, Cc, ws, dgamma, dbeta, dy, dz_out=None, dA2=None, ng=1, param_gstride=0,
           grad_gstride=0, reverse=False, mask_bits=None, policy_groups=None):
    """gamma / dgamma / dbeta: pointers (ints) or tensors, + g*param_gstride / g*grad_gstride per group;
    mask_bits: the bit plane written by bn_apply (mask_out), used instead of the bf16 plane mask_hi"""
    def p(t):
        return t if isinstance(t, int) or t is None else t.data_ptr()

    a = L.BnBwdArgs()
    a.dA, a.dA2, a.mask_hi, a.y, a.mean, a.rstd, a.gamma = dA.data_ptr(), L.ptr(dA2), L.ptr(mask_hi), y.data_ptr(), \
        mean.data_ptr(), rstd.data_ptr(), p(gamma)
    a.P, a.C, a.ws = P, Cc, ws.data_ptr()
    a.dgamma, a.dbeta, a.dy_bf16 = p(dgamma), p(dbeta), dy.data_ptr()
    a.dz_out = L.ptr(dz_out)
    a.ng, a.param_gstride, a.grad_gstride = ng, param_gstride, grad_gstride
    a.policy_groups, a.reverse = int(policy_groups or POLICY_GROUPS), int(reverse)
    a.mask_bits = L.ptr(mask_bits)
    # distinct tensors: dA, y (+ mask) in; dy (+ dz) out -- each counted once although the two launches read twice
    mask_bytes = 0.125 if mask_bits is not None else (2.0 if mask_hi is not None else 0.0)
    per_elem = 4.0 + 4.0 + mask_bytes + 2.0 + (4.0 if dz_out is not None else 0.0) + (4.0 if dA2 is not None else 0.0)
    _call("bn_bwd", per_elem * P * Cc * ng, "byte", "fb_bn_bwd", C.byref(a), label=f"bn_bwd P{P} C{Cc} {per_elem:.0f}B/elem")


def avgpool2_fwd(in_hi, in_lo, n, h, w, c, out_hi, out_lo):
    planes = 1 + (in_lo is not None)
    _call("pool", n * h * w * c * 2.0 * planes * 1.25, "byte", "fb_avgpool2_fwd", in_hi.data_ptr(), L.ptr(in_lo), n, h, w,
          c, out_hi.data_ptr(), L.ptr(out_lo))


def avgpool2_bwd(dP, n, h, w, c, dX, accumulate=False):
    _call("pool", n * h * w * c * 4.0 * (1.25 + (1.0 if accumulate else 0.0)), "byte", "fb_avgpool2_bwd", dP.data_ptr(),
          n, h, w, c, dX.data_ptr(), int(accumulate))


def head_ws_floats(n_total, c):
    return n_total * (c + 32)


def head_fwd_bwd(a_hi, a_lo, n, hw, c, fc_w, fc_b, labels, classes, smoothing, ws, scal, loss_base, correct_base, d_fcw,
                 d_fcb, dA, ng=1, param_gstride=0, grad_gstride=0):
    """fc_w / fc_b / d_fcw / d_fcb: pointers (ints), + g*param_gstride / g*grad_gstride per group"""
    planes = 1 + (a_lo is not None)
    nbytes = ng * n * hw * c * (2.0 * planes + 4.0)
    _call("head", nbytes, "byte", "fb_head_fwd_bwd", a_hi.data_ptr(), L.ptr(a_lo), n, hw, c, fc_w, fc_b,
          labels.data_ptr(), classes, smoothing, ws.data_ptr(), scal.data_ptr(), loss_base, correct_base, d_fcw, d_fcb,
          dA.data_ptr(), ng, param_gstride, grad_gstride)


def flat_sqnorm(x, n, ws, scal, slot_base, ng=1, gstride=0, y=None, a=1.0, b=0.0, norms_out=None, cursor=None, eps_mode=0,
                bs=0.0, eps=0.0, eps_base=0):
    _call("flat", 4.0 * n * ng * (1 + (y is not None)), "byte", "fb_flat_sqnorm", x.data_ptr(), gstride, L.ptr(y), a, b,
          n, ng, ws.data_ptr(), scal.data_ptr(), slot_base, L.ptr(norms_out), L.ptr(cursor), eps_mode, bs, eps, eps_base)


def perturb_ranges(theta, grad, gstride, pre, ranges, n_ranges, total, bs, acc, scale, scal, eps_base, theta_p,
                   theta_p_gstride, ng):
    _call("flat", 4.0 * total * ng * (3 + (pre is not None)), "byte", "fb_perturb_ranges", theta.data_ptr(),
          grad.data_ptr(), gstride, L.ptr(pre), ranges.data_ptr(), n_ranges, total, bs, acc, scale, scal.data_ptr(),
          eps_base, theta_p.data_ptr(), theta_p_gstride, ng)


def fd_combine(grad, g_plus, g_minus, gstride, avg, n, ng, scal, eps_base, cf_slot, cursor, write_g):
    per = ng * (2 + (g_minus is not None) + (1 if write_g else 0)) + (2 if avg is not None else 0)
    _call("flat", 4.0 * n * per, "byte", "fb_fd_combine", grad.data_ptr(), g_plus.data_ptr(), L.ptr(g_minus), gstride,
          L.ptr(avg), n, ng, scal.data_ptr(), eps_base, cf_slot, L.ptr(cursor), int(write_g))


def mean_accumulate(grad, gstride, avg, n, ng, cursor, scal=None, norm_base=0, clip=0.0, clipped_slot=0):
    _call("flat", 4.0 * n * (ng + 2), "byte", "fb_mean_accumulate", grad.data_ptr(), gstride, avg.data_ptr(), n, ng,
          L.ptr(cursor), L.ptr(scal), norm_base, clip, clipped_slot)


def group_finish(cursor, ng, scal, loss_slot, correct_slot, loss_base, correct_base, cursor_step=None, totals=None):
    """totals: the scalar block that holds the loss / accuracy sums (shared by all lanes; default: scal itself)"""
    _call("misc", 8.0 * ng + 16.0, "byte", "fb_group_finish", cursor.data_ptr(), ng,
          ng if cursor_step is None else cursor_step, scal.data_ptr(), (scal if totals is None else totals).data_ptr(),
          loss_slot, correct_slot, loss_base, correct_base)


def flat_scale(x, n, alpha):
    _call("flat", 8.0 * n, "byte", "fb_flat_scale", x.data_ptr(), n, alpha)


def flat_relayout(src, dst, n, table, n_entries, total_blocks, to_native):
    _call("flat", 8.0 * n, "byte", "fb_flat_relayout", src.data_ptr(), dst.data_ptr(), n, L.ptr(table), n_entries,
          total_blocks, int(to_native))


def sgd_step(theta, grad, buf, n, scal, norm_slot, clip, lr, momentum, dampening, wd, nesterov, first, write_grad, ws,
             param_norm_slot):
    per = 3 + (2 if buf is not None else 0) + (1 if write_grad else 0)
    _call("flat", 4.0 * n * per, "byte", "fb_sgd_step", theta.data_ptr(), grad.data_ptr(), L.ptr(buf), n, scal.data_ptr(),
          norm_slot, clip, lr, momentum, dampening, wd, int(nesterov), int(first), int(write_grad), ws.data_ptr(),
          param_norm_slot)
