"""Host-side operator layer over the C ABI: builds TMA descriptors / tap tables for each convolution and wraps the
layer kernels.  Everything here only prepares arguments; all arithmetic happens in the CUDA library (lib.py).

Layouts: activations NHWC; a "split" activation is a pair of bf16 planes (hi, lo) with x ~= hi + lo (lo is None in
plain-bf16 mode); conv outputs / activation gradients fp32 NHWC; dY bf16 NHWC; GEMM weights bf16 hi/lo as
wf[cout][tap*cin + ci] and wd[cin][tap*cout + co].
"""
import ctypes as C
import os

import torch

from . import lib as L

NUM_SMS = 148

# Optional per-launch timing for bench.py's roofline: set PROFILE = [] to collect
# (family, algorithmic work, unit, start event, end event) around every wrapped launch on the current stream.
PROFILE = None
LAUNCHES = {"count": 0}  # kernels launched through this module (bench.py's gpu_launches claim)
_KERNELS_PER_CALL = {"fb_weight_prep_multi": 1, "fb_conv_gemm": 1, "fb_conv3x3": 1, "fb_conv_wgrad": 1, "fb_wgrad_finalize": 1, "fb_weight_prep": 1,
                     "fb_stem_im2col": 1, "fb_stem_im2col_u8aug": 1, "fb_bn_fwd_fused": 1, "fb_bn_bwd_fused": 1, "fb_bn_stats": 2, "fb_bn_apply": 1, "fb_bn_bwd": 3, "fb_avgpool2_fwd": 1,
                     "fb_avgpool2_bwd": 1, "fb_head_fwd_bwd": 3, "fb_flat_sqnorm": 2, "fb_fd_perturb": 1,
                     "fb_fd_combine": 1, "fb_mean_accumulate": 1, "fb_cursor_add": 1, "fb_flat_scale": 1, "fb_sgd_step": 2,
                     "fb_flat_sqnorm_axpby": 2, "fb_fd_perturb_ex": 1, "fb_fd_combine_ex": 1,
                     "fb_mean_accumulate_clip": 1}


def _call(family, work, unit, name, *args):
    LAUNCHES["count"] += _KERNELS_PER_CALL[name]
    if PROFILE is None:
        L.call(name, *args)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.call(name, *args)
    e1.record()
    PROFILE.append((family, work, unit, e0, e1))


def pixel_tile(h, w):
    """128-pixel TMA box (tile_w, tile_h, tile_n) over an [n, h, w] pixel grid; the box always spans full rows."""
    if w > 128 or 128 % w != 0:
        raise RuntimeError(f"unsupported feature-map width {w} (no fallback path)")
    tile_h = min(h, 128 // w)
    if h % tile_h != 0 or 128 % (w * tile_h) != 0:
        raise RuntimeError(f"unsupported feature-map size {h}x{w}")
    return w, tile_h, 128 // (w * tile_h)


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
    """One launch of fb_conv_gemm with frozen arguments (descriptors are encoded once)."""

    def __init__(self, a_maps, b_maps, n_phases, a_planes, b_planes, taps, cblocks, tile, grid_h, grid_n, n_total, out,
                 out_off, out_strides, accumulate, n_tile, groups=None):
        """groups: optional [(tap0, n_taps, out_off_elements)] -- tap groups that run over the same pixel grid and
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
        args.grid_h, args.grid_n = grid_h, grid_n
        args.n_total, args.n_tile = n_total, n_tile
        self.out = out
        args.out = out.data_ptr() + out_off * 4
        args.out_sn, args.out_sh, args.out_sw = out_strides
        args.accumulate = int(accumulate)
        if groups:
            if len(groups) > 4:
                raise RuntimeError("at most 4 tap groups")
            args.n_groups = len(groups)
            for i, (tap0, n_taps, off) in enumerate(groups):
                args.groups[i] = L.TapGroup(tap0, n_taps, off)
        self.args = args
        self.flops = 0.0  # algorithmic FLOPs of this launch (set by Conv2dPlan)

    def __call__(self):
        _call("conv_gemm", self.flops, "flop", "fb_conv_gemm", C.byref(self.args))


FABRIC_BYTES_PER_CLK = 43.0  # measured L2 -> SM operand bandwidth per SM


def halo_geometry(n, h, w, k, stride, n_total, a_planes=2, b_planes=2):
    """Tile geometry (imgs, halves, n_tile) of the haloed-box kernel for a 3x3 / stride-1 conv, or None.

    imgs == 1: tiles are whole image rows (>= 1024 bytes per row => w >= 8 with 128 % w == 0).  Small maps
    (imgs * w * h == 128): a 128-pixel half is `imgs` whole images with interleaved rows.  halves and the N tile are
    chosen by a cost model: waves of the persistent grid x max(tensor-pipe clocks, L2->SM operand bytes / 43 B/clk)
    per (tap, 64-channel block)."""
    if os.environ.get("FB_HALO", "0") != "1" or k != 3 or stride != 1:
        # opt-in: on B200 the generic per-tap kernel (deeper operand pipeline) is faster on every ResNet shape although
        # it moves 1.5-2.2x more operand bytes (tools/conv_geom_sweep.py)
        return None
    if w in (16, 32, 64, 128) and h % (128 // w) == 0:
        imgs = 1
    elif w >= 4 and h * w <= 64 and 128 % (h * w) == 0 and (128 // (h * w) * w) % 8 == 0:
        imgs = 128 // (h * w)
    else:
        return None
    th = 128 // (imgs * w)
    best = None
    for halves in (2, 1):
        if imgs == 1:
            if h % (halves * th) != 0:
                continue
            m_tiles = n * (h // (halves * th))
            box_bytes = (halves * th + 2) * w * 128
        else:
            if n % (halves * imgs) != 0:
                continue
            m_tiles = n // (halves * imgs)
            box_bytes = halves * (th + 2) * imgs * w * 128
        for nt in (128, 64):
            if n_total % nt != 0:
                continue
            if nt == 64 and b_planes == 2:
                mma = a_planes * 64
            else:
                combos = 3 if (a_planes == 2 and b_planes == 2) else a_planes * b_planes
                mma = combos * 64
            mma *= 4 * halves  # four K=16 steps per 64-channel block, per half
            fabric = (a_planes * box_bytes / 3.0 + b_planes * nt * 128) / FABRIC_BYTES_PER_CLK
            a_smem = 2 * a_planes * box_bytes  # two A stages
            if a_smem + 2 * b_planes * nt * 128 > 210 * 1024:
                continue
            waves = -(-(m_tiles * (n_total // nt)) // NUM_SMS)
            cost = waves * max(mma, fabric)
            if best is None or cost < best[0]:
                best = (cost, imgs, halves, nt)
    return None if best is None else best[1:]


class Conv3x3:
    """One launch of fb_conv3x3 (haloed A boxes; 1 or 2 128-pixel halves per tile) with frozen arguments."""

    def __init__(self, planes_a, planes_b, n, h, w, c_k, n_total, b_k0, out, accumulate, geom=None):
        """planes_a: list of NHWC bf16 planes [n,h,w,c_k]; planes_b: list of weight matrices [n_total][9*c_k];
        geom = (imgs, halves, n_tile) from halo_geometry."""
        if geom is None:
            geom = halo_geometry(n, h, w, 3, 1, n_total, len(planes_a), len(planes_b))
        if geom is None:
            raise RuntimeError(f"fb_conv3x3 does not support {n} images of {h}x{w}")
        imgs, halves, n_tile = geom
        th = 128 // (imgs * w)
        self.a_maps = MapSet(len(planes_a))
        for i, t in enumerate(planes_a):
            if imgs == 1:
                encode_act(self.a_maps, i, t, n, h, w, c_k, (w, halves * th + 2, 1))
            else:
                # dims (C, W, N, H): the box holds rows -1..h of `imgs` images, image-interleaved per row
                assert t.dtype == torch.bfloat16 and t.is_contiguous()
                L.check(L.load().fb_tmap_encode_act4d(self.a_maps.slot(i), t.data_ptr(), c_k, w, n, h, c_k,
                                                      h * w * c_k, w * c_k, 64, w, imgs, h + 2),
                        "fb_tmap_encode_act4d")
                self.a_maps.keep.append(t)
        self.b_maps = MapSet(len(planes_b))
        for i, t in enumerate(planes_b):
            encode_mat(self.b_maps, i, t, 9 * c_k, n_total, n_tile)
        a = L.Conv3x3Args()
        a.host_a_maps, a.host_b_maps = self.a_maps.addr, self.b_maps.addr
        a.a_planes, a.b_planes = len(planes_a), len(planes_b)
        for i in range(3):
            for j in range(3):
                a.b_k0[i][j] = b_k0[i][j]
        a.cblocks = c_k // 64
        a.w, a.h, a.n = w, h, n
        a.n_total, a.n_tile = n_total, n_tile
        a.imgs, a.halves = imgs, halves
        self.m_tiles = n * (h // (halves * th)) if imgs == 1 else n // (halves * imgs)
        self.out = out
        a.out = out.data_ptr()
        a.out_sn, a.out_sh, a.out_sw = h * w * n_total, w * n_total, n_total
        a.accumulate = int(accumulate)
        self.args = a
        self.flops = 0.0

    def __call__(self):
        _call("conv_gemm", self.flops, "flop", "fb_conv3x3", C.byref(self.args))


class Conv2dPlan:
    """Forward / dgrad / wgrad launches of one bias-free Conv2d (k in {1,3}, stride in {1,2}, pad (k-1)/2) on NHWC data.

    x_hi/x_lo : [n,h,w,cin] bf16 planes (conv input; for the stem: the im2col patches with cin=64, k=1)
    y         : [n,ho,wo,cout] fp32 (forward output)
    dy        : [n,ho,wo,cout] bf16 (output gradient)
    dx        : [n,h,w,cin] fp32 (input gradient) or None if no dgrad is needed
    wf/wd     : weight operand matrices (hi, lo) made by fb_weight_prep
    """

    def __init__(self, n, h, w, cin, cout, k, stride, x_hi, x_lo, y, dy, dx, wf_hi, wf_lo, wd_hi, wd_lo, partial,
                 dx_accumulate=False, split=True, alg_k=None, fuse_stats=False, dgrad_bn=None):
        assert k in (1, 3) and stride in (1, 2) and cin % 64 == 0 and cout % 64 == 0
        assert not (k == 1 and stride == 2)
        self.n, self.h, self.w, self.cin, self.cout, self.k, self.stride = n, h, w, cin, cout, k, stride
        ho, wo = h // stride, w // stride
        self.ho, self.wo = ho, wo
        taps = k * k
        self.taps = taps
        # algorithmic FLOPs of one forward (= one dgrad = one wgrad): 2 * pixels * cout * K, K = taps*cin (27 for the stem)
        self.alg_flops = 2.0 * n * ho * wo * cout * (alg_k or taps * cin)
        planes = 2 if (split and x_lo is not None) else 1
        wplanes = 2 if (split and wf_lo is not None) else 1
        tile = pixel_tile(ho, wo)
        cb_in, cb_out = cin // 64, cout // 64

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

        # ---- forward
        m_tiles = n * (ho // tile[1]) if tile[2] == 1 else -(-n // tile[2])
        n_tile = choose_n_tile(m_tiles, cout, planes, wplanes)
        bs = MapSet(wplanes)
        for pl, t in enumerate((wf_hi, wf_lo)[:wplanes]):
            encode_mat(bs, pl, t, taps * cin, cout, n_tile)
        if planes != wplanes:
            raise RuntimeError("activation and weight operands must both be split or both be plain bf16")
        ftaps = []
        for kh in range(k):
            for kw in range(k):
                phase, dh, dw = tap_geom(kh, kw)
                ftaps.append((phase, dh, dw, (kh * k + kw) * cin))
        halo = halo_geometry(n, h, w, k, stride, cout, planes, wplanes)
        if halo:
            # b_k0[dw+1][dh+1]: forward tap (kh, kw) reads input pixel (h + kh - 1, w + kw - 1)
            fk0 = [[(dhi * 3 + dwi) * cin for dhi in range(3)] for dwi in range(3)]
            self.fwd = Conv3x3([x_hi, x_lo][:planes], [wf_hi, wf_lo][:wplanes], n, h, w, cin, cout, fk0, y, False,
                               halo)
        else:
            self.fwd = ConvGemm(xs, bs, nph, planes, wplanes, ftaps, cb_in, tile, ho, n, cout, y, 0,
                                (ho * wo * cout, wo * cout, cout), False, n_tile)
        self.fwd.flops = self.alg_flops
        # BatchNorm statistics fused into the forward epilogue: per-CTA partial rows [rows][2][cout]
        self.stats = None
        if fuse_stats:
            a = self.fwd.args
            if halo:
                m_t, n_t = self.fwd.m_tiles, cout // a.n_tile
            else:
                m_t, n_t = m_tiles, cout // a.n_tile
            rows = L.load().fb_conv_stats_rows(m_t, n_t)
            buf = torch.zeros(rows, 2, cout, device=y.device)
            a.stats_out = buf.data_ptr()
            self.stats = (buf, rows)

        # ---- dgrad.  dgrad_bn = (y, mask_hi, mean, rstd) of the BatchNorm(+ReLU) whose upstream gradient is `dx` and has
        # no other producer: the dgrad epilogue then also reduces that BatchNorm's backward statistics (dgrad_stats).
        self.dgrads = []
        self.dgrad_stats = None
        if dx is not None:
            dys = MapSet(1)
            encode_act(dys, 0, dy, n, ho, wo, cout, tile)
            grouped = stride == 2 and os.environ.get("FB_S2_DGRAD_GROUPS", "1") == "1"
            m_tiles_d = m_tiles * (4 if grouped else 1)  # the four output phases of a stride-2 dgrad share one launch
            n_tile_d = choose_n_tile(m_tiles_d, cin, 1, wplanes)
            ds = MapSet(wplanes)
            for pl, t in enumerate((wd_hi, wd_lo)[:wplanes]):
                encode_mat(ds, pl, t, taps * cout, cin, n_tile_d)
            halo_d = halo_geometry(n, h, w, k, stride, cin, 1, wplanes)
            if halo_d:
                # dgrad tap (kh, kw) reads dY pixel (h + 1 - kh, w + 1 - kw): dh = 1 - kh, dw = 1 - kw
                dk0 = [[((2 - dhi) * 3 + (2 - dwi)) * cout for dhi in range(3)] for dwi in range(3)]
                self.dgrads.append(Conv3x3([dy], [wd_hi, wd_lo][:wplanes], n, h, w, cout, cin, dk0, dx, dx_accumulate,
                                           halo_d))
                self.dgrads[-1].flops = self.alg_flops
            elif stride == 1:
                dtaps = []
                for kh in range(k):
                    for kw in range(k):
                        dh, dw = (1 - kh, 1 - kw) if k == 3 else (0, 0)
                        dtaps.append((0, dh, dw, (kh * k + kw) * cout))
                self.dgrads.append(ConvGemm(dys, ds, 1, 1, wplanes, dtaps, cb_out, tile, ho, n, cin, dx, 0,
                                            (h * w * cin, w * cin, cin), dx_accumulate, n_tile_d))
                self.dgrads[-1].flops = self.alg_flops
                if dgrad_bn is not None and not dx_accumulate:
                    by, bmask, bmean, brstd = dgrad_bn
                    a = self.dgrads[-1].args
                    rows = L.load().fb_conv_stats_rows(m_tiles_d, cin // n_tile_d)
                    buf = torch.zeros(rows, 2, cin, device=dx.device)
                    a.stats_out = buf.data_ptr()
                    a.bwd_y, a.bwd_mask = by.data_ptr(), L.ptr(bmask)
                    a.bwd_mean, a.bwd_rstd = bmean.data_ptr(), brstd.data_ptr()
                    self.dgrad_stats = (buf, rows)
                    self._dgrad_bn_keep = dgrad_bn
            else:
                # stride 2: output pixel (2i+ph, 2j+pw) gathers taps kh with (ph + 1 - kh) even: ho = i + (ph+1-kh)/2
                def taps_for(par):
                    return [(1, 0)] if par == 0 else [(0, 1), (2, 0)]  # (k index, shift in the dY grid)

                # one launch: four tap groups = the four output phases (1 + 2 + 2 + 4 taps)
                dtaps, groups = [], []
                for ph in range(2):
                    for pw in range(2):
                        tap0 = len(dtaps)
                        for kh, dh in taps_for(ph):
                            for kw, dw in taps_for(pw):
                                dtaps.append((0, dh, dw, (kh * 3 + kw) * cout))
                        groups.append((tap0, len(dtaps) - tap0, (ph * w + pw) * cin))
                if grouped:
                    self.dgrads.append(ConvGemm(dys, ds, 1, 1, wplanes, dtaps, cb_out, tile, ho, n, cin, dx, 0,
                                                (h * w * cin, 2 * w * cin, 2 * cin), dx_accumulate, n_tile_d,
                                                groups=groups))
                    self.dgrads[-1].flops = self.alg_flops
                else:
                    for tap0, cnt, off in groups:
                        self.dgrads.append(ConvGemm(dys, ds, 1, 1, wplanes, dtaps[tap0:tap0 + cnt], cb_out, tile, ho, n,
                                                    cin, dx, off, (h * w * cin, 2 * w * cin, 2 * cin), dx_accumulate,
                                                    n_tile_d))
                        self.dgrads[-1].flops = self.alg_flops * cnt / 9.0
            self.dy_maps_d = dys

        # ---- wgrad
        wa = L.WgradArgs()
        dym = MapSet(1)
        encode_act(dym, 0, dy, n, ho, wo, cout, tile)
        self.dy_map_w = dym
        n_slots = taps * cb_in
        co_tiles = -(-cout // 128)
        n_pixblocks = m_tiles
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
        groups = -(-n_slots // spc)
        splits = max(1, min(n_pixblocks, NUM_SMS // (co_tiles * groups)))
        wa.slots_per_cta = spc
        wa.cout, wa.cin = cout, cin
        wa.tile_w, wa.tile_h, wa.tile_n = tile
        wa.grid_h, wa.grid_n = ho, n
        wa.splits = splits
        need = splits * cout * taps * cin
        if partial.numel() < need:
            raise RuntimeError(f"wgrad workspace too small: {partial.numel()} < {need}")
        self.partial = partial
        wa.partial = partial.data_ptr()
        self.wargs = wa
        self.splits = splits

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

    @staticmethod
    def partial_elems(n, h, w, cin, cout, k, stride, planes=2):
        """Upper bound of the split-K workspace (fp32 elements) over both operand modes."""
        ho, wo = h // stride, w // stride
        tile = pixel_tile(ho, wo)
        m_tiles = n * (ho // tile[1]) if tile[2] == 1 else -(-n // tile[2])
        taps = k * k
        n_slots = taps * (cin // 64)
        worst = 0
        for pl in (1, 2):
            spc = Conv2dPlan._slots_per_cta(n_slots, pl)
            splits = max(1, min(m_tiles, NUM_SMS // ((-(-cout // 128)) * (-(-n_slots // spc)))))
            worst = max(worst, splits)
            if k == 3 and stride == 1 and tile[2] == 1:  # haloed variant: 3 slots per CTA
                worst = max(worst, max(1, min(m_tiles, NUM_SMS // ((-(-cout // 128)) * (-(-n_slots // 3))))))
        return worst * cout * taps * cin

    def forward(self):
        self.fwd()

    def dgrad(self):
        for d in self.dgrads:
            d()

    def wgrad(self, g_oihw, cin_real=None, mode=0):
        """partial sums -> fixed-order reduction -> g (OIHW fp32 view of the flat gradient buffer)."""
        _call("conv_wgrad", self.alg_flops, "flop", "fb_conv_wgrad", C.byref(self.wargs))
        cin_real = cin_real or self.cin
        taps = self.taps if mode == 0 else 9
        nbytes = 4.0 * self.cout * (self.splits * self.taps * self.cin + cin_real * taps)
        _call("wgrad_finalize", nbytes, "byte", "fb_wgrad_finalize", self.partial.data_ptr(), self.splits, self.cout,
              cin_real, taps, self.cin, mode, g_oihw.data_ptr())


# ---- thin wrappers of the layer kernels ------------------------------------------------------------------------------

def weight_prep(w_oihw, cout, cin, taps, wf_hi, wf_lo, wd_hi=None, wd_lo=None):
    planes = (1 + (wf_lo is not None)) * (1 + (wd_hi is not None))
    _call("weight_prep", cout * cin * taps * (4.0 + 2.0 * planes), "byte", "fb_weight_prep", w_oihw.data_ptr(), cout, cin,
          taps, wf_hi.data_ptr(), L.ptr(wf_lo), wf_hi.stride(0),
           L.ptr(wd_hi), L.ptr(wd_lo), wd_hi.stride(0) if wd_hi is not None else 0)


class WeightPrepTable:
    """Device table for fb_weight_prep_multi: entries = (w_offset, cout, cin, taps, wf_hi, wf_lo, wd_hi, wd_lo)."""

    def __init__(self, entries, device):
        arr = (L.WprepEntry * len(entries))()
        block, nbytes = 0, 0.0
        for i, (off, cout, cin, taps, wf_hi, wf_lo, wd_hi, wd_lo) in enumerate(entries):
            nb = (cout // 32) * (cin // 32) if cin % 32 == 0 else 4
            arr[i] = L.WprepEntry(off, cout, cin, taps, block, nb, 0, wf_hi.data_ptr(), L.ptr(wf_lo), L.ptr(wd_hi),
                                  L.ptr(wd_lo), wf_hi.stride(0), wd_hi.stride(0) if wd_hi is not None else 0)
            block += nb
            planes = (1 + (wf_lo is not None)) * (1 + (wd_hi is not None))
            nbytes += cout * cin * taps * (4.0 + 2.0 * planes)
        self.keep = entries
        self.n, self.blocks, self.nbytes = len(entries), block, nbytes
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self.table = raw.to(device)

    def __call__(self, theta):
        _call("weight_prep", self.nbytes, "byte", "fb_weight_prep_multi", theta.data_ptr(), self.table.data_ptr(), self.n,
              self.blocks)


def stem_im2col(x, labels, perm, cursor, first, n, p_hi, p_lo, labels_out):
    _call("stem_im2col", n * (3072 * 4.0 + 1024 * 64 * 2.0 * (1 + (p_lo is not None))), "byte", "fb_stem_im2col",
          x.data_ptr(), L.ptr(labels), L.ptr(perm), L.ptr(cursor), first, n, p_hi.data_ptr(),
           L.ptr(p_lo), L.ptr(labels_out))


def stem_im2col_u8aug(x_u8, labels, perm, cursor, first, n, aug, mean, std, p_hi, p_lo, labels_out):
    """x_u8: [N,32,32,3] uint8 on the device; aug: int8 [.,4] (dx, dy, flip, 0) per position or None; mean/std: 3 floats"""
    m = (L.f32 * 3)(*[float(v) for v in mean])
    sd = (L.f32 * 3)(*[float(v) for v in std])
    _call("stem_im2col", n * (3072.0 + 1024 * 64 * 2.0 * (1 + (p_lo is not None))), "byte", "fb_stem_im2col_u8aug",
          x_u8.data_ptr(), L.ptr(labels), L.ptr(perm), L.ptr(cursor), first, n, L.ptr(aug), m, sd, p_hi.data_ptr(),
          L.ptr(p_lo), L.ptr(labels_out))


def bn_stats(y, P, Cc, ws, mean, rstd, running_mean, running_var, momentum=0.1, eps=1e-5):
    _call("bn_stats", 4.0 * P * Cc, "byte", "fb_bn_stats", y.data_ptr(), P, Cc, ws.data_ptr(), mean.data_ptr(), rstd.data_ptr(), L.ptr(running_mean),
           L.ptr(running_var), momentum, eps)


def bn_apply(y, mean, rstd, gamma, beta, P, Cc, out_hi, out_lo, relu=True, second=None, res=None):
    a = L.BnApplyArgs()
    a.y, a.mean, a.rstd, a.gamma, a.beta = (t.data_ptr() for t in (y, mean, rstd, gamma, beta))
    if second is not None:
        a.y2, a.mean2, a.rstd2, a.gamma2, a.beta2 = (t.data_ptr() for t in second)
    if res is not None:
        a.res_hi, a.res_lo = res[0].data_ptr(), L.ptr(res[1])
    a.relu, a.P, a.C = int(relu), P, Cc
    a.out_hi, a.out_lo = out_hi.data_ptr(), L.ptr(out_lo)
    planes = 1 + (out_lo is not None)
    per_elem = 4.0 + (4.0 if second is not None else 0.0) + (2.0 * planes if res is not None else 0.0) + 2.0 * planes
    _call("bn_apply", per_elem * P * Cc, "byte", "fb_bn_apply", C.byref(a))


def bn_bwd(dA, mask_hi, y, mean, rstd, gamma, P, Cc, ws, dgamma, dbeta, dy, dz_out=None, dz_accumulate=False, dA2=None):
    a = L.BnBwdArgs()
    a.dA, a.dA2, a.mask_hi, a.y, a.mean, a.rstd, a.gamma = dA.data_ptr(), L.ptr(dA2), L.ptr(mask_hi), y.data_ptr(), \
        mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr()
    a.P, a.C, a.ws = P, Cc, ws.data_ptr()
    a.dgamma, a.dbeta, a.dy_bf16 = dgamma.data_ptr(), dbeta.data_ptr(), dy.data_ptr()
    a.dz_out, a.dz_accumulate = L.ptr(dz_out), int(dz_accumulate)
    # distinct tensors: dA, y (+ mask) in; dy (+ dz) out -- each counted once although the two-phase kernel reads twice
    per_elem = 4.0 + 4.0 + (2.0 if mask_hi is not None else 0.0) + 2.0 + (4.0 if dz_out is not None else 0.0) + \
        (4.0 if dA2 is not None else 0.0)
    _call("bn_bwd", per_elem * P * Cc, "byte", "fb_bn_bwd", C.byref(a))


def bn_fwd_fused(y, mean, rstd, gamma, beta, P, Cc, out_hi, out_lo, ws, running=None, relu=True, second=None, res=None,
                 momentum=0.1, eps=1e-5, stats=None, stats2=None):
    """Train-mode BatchNorm statistics + normalise (+ second normalised branch / residual) + ReLU in one launch.
    second = (y2, mean2, rstd2, gamma2, beta2, running_mean2, running_var2)."""
    a = L.BnApplyArgs()
    a.y, a.mean, a.rstd, a.gamma, a.beta = (t.data_ptr() for t in (y, mean, rstd, gamma, beta))
    m2 = r2 = rm2 = rv2 = None
    if second is not None:
        y2, m2, r2, g2, b2, rm2, rv2 = second
        a.y2, a.mean2, a.rstd2, a.gamma2, a.beta2 = (t.data_ptr() for t in (y2, m2, r2, g2, b2))
    if res is not None:
        a.res_hi, a.res_lo = res[0].data_ptr(), L.ptr(res[1])
    a.relu, a.P, a.C = int(relu), P, Cc
    a.out_hi, a.out_lo = out_hi.data_ptr(), L.ptr(out_lo)
    planes = 1 + (out_lo is not None)
    per_elem = 4.0 + (4.0 if second is not None else 0.0) + (2.0 * planes if res is not None else 0.0) + 2.0 * planes
    rm, rv = running if running is not None else (None, None)
    # stats = (tensor [rows][2][C], rows) produced by the convolution's epilogue (Conv2dPlan.stats)
    st, st_rows = (stats[0].data_ptr(), stats[1]) if stats is not None else (None, 0)
    st2, st_rows2 = (stats2[0].data_ptr(), stats2[1]) if stats2 is not None else (None, 0)
    _call("bn_fwd", per_elem * P * Cc, "byte", "fb_bn_fwd_fused", C.byref(a), L.ptr(m2), L.ptr(r2), L.ptr(rm), L.ptr(rv),
          L.ptr(rm2), L.ptr(rv2), momentum, eps, ws.data_ptr(), st, st_rows, st2, st_rows2)


def bn_bwd_fused(dA, mask_hi, y, mean, rstd, gamma, P, Cc, ws, dgamma, dbeta, dy, dz_out=None, dA2=None, stats=None):
    """stats = (buffer, rows): partial sums written by the dgrad that produced dA (Conv2dPlan.dgrad_stats)."""
    a = L.BnBwdArgs()
    a.dA, a.dA2, a.mask_hi, a.y, a.mean, a.rstd, a.gamma = dA.data_ptr(), L.ptr(dA2), L.ptr(mask_hi), y.data_ptr(), \
        mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr()
    a.P, a.C, a.ws = P, Cc, ws.data_ptr()
    a.dgamma, a.dbeta, a.dy_bf16 = dgamma.data_ptr(), dbeta.data_ptr(), dy.data_ptr()
    a.dz_out, a.dz_accumulate = L.ptr(dz_out), 0
    if stats is not None:
        a.stats, a.stats_rows = stats[0].data_ptr(), stats[1]
    per_elem = 4.0 + 4.0 + (2.0 if mask_hi is not None else 0.0) + 2.0 + (4.0 if dz_out is not None else 0.0) + \
        (4.0 if dA2 is not None else 0.0)
    _call("bn_bwd", per_elem * P * Cc, "byte", "fb_bn_bwd_fused", C.byref(a))


def avgpool2_fwd(in_hi, in_lo, n, h, w, c, out_hi, out_lo):
    _call("misc", 0.0, "byte", "fb_avgpool2_fwd", in_hi.data_ptr(), L.ptr(in_lo), n, h, w, c, out_hi.data_ptr(), L.ptr(out_lo))


def avgpool2_bwd(dP, n, h, w, c, dX, accumulate=False):
    _call("misc", 0.0, "byte", "fb_avgpool2_bwd", dP.data_ptr(), n, h, w, c, dX.data_ptr(), int(accumulate))


def head_fwd_bwd(a_hi, a_lo, n, hw, c, fc_w, fc_b, labels, classes, smoothing, ws, scal, loss_slot, correct_slot, d_fcw,
                 d_fcb, dA):
    _call("misc", 0.0, "byte", "fb_head_fwd_bwd", a_hi.data_ptr(), L.ptr(a_lo), n, hw, c, fc_w.data_ptr(), fc_b.data_ptr(),
           labels.data_ptr(), classes, smoothing, ws.data_ptr(), scal.data_ptr(), loss_slot, correct_slot,
           d_fcw.data_ptr(), d_fcb.data_ptr(), dA.data_ptr())


def flat_sqnorm(x, n, ws, scal, slot, norms_out=None, cursor=None):
    _call("misc", 0.0, "byte", "fb_flat_sqnorm", x.data_ptr(), n, ws.data_ptr(), scal.data_ptr(), slot, L.ptr(norms_out), L.ptr(cursor))


def fd_perturb(theta, g, n, bs, eps, scal, sq_slot, eps_slot, theta_p):
    _call("misc", 0.0, "byte", "fb_fd_perturb", theta.data_ptr(), g.data_ptr(), n, bs, eps, scal.data_ptr(), sq_slot, eps_slot,
           theta_p.data_ptr())


def fd_combine(g, g2, avg, n, scal, eps_slot, cf, cursor, count0, write_g, cf_slot=-1):
    _call("misc", 0.0, "byte", "fb_fd_combine", g.data_ptr(), g2.data_ptr(), L.ptr(avg), n, scal.data_ptr(), eps_slot, cf, cf_slot,
           L.ptr(cursor), count0, int(write_g))


def mean_accumulate(g, avg, n, cursor, count0):
    _call("misc", 0.0, "byte", "fb_mean_accumulate", g.data_ptr(), avg.data_ptr(), n, L.ptr(cursor), count0)


def cursor_add(cursor, delta):
    _call("misc", 0.0, "byte", "fb_cursor_add", cursor.data_ptr(), delta)


def flat_scale(x, n, alpha):
    _call("misc", 0.0, "byte", "fb_flat_scale", x.data_ptr(), n, alpha)


def sgd_step(theta, grad, buf, n, scal, norm_slot, clip, lr, momentum, dampening, wd, nesterov, first, write_grad, ws,
             param_norm_slot):
    _call("misc", 0.0, "byte", "fb_sgd_step", theta.data_ptr(), grad.data_ptr(), L.ptr(buf), n, scal.data_ptr(), norm_slot,
          clip, lr, momentum, dampening, wd, int(nesterov), int(first), int(write_grad), ws.data_ptr(), param_norm_slot)


def flat_sqnorm_axpby(x, y, a, b, n, ws, scal, slot):
    _call("misc", 0.0, "byte", "fb_flat_sqnorm_axpby", x.data_ptr(), L.ptr(y), a, b, n, ws.data_ptr(), scal.data_ptr(), slot)


def fd_perturb_ex(theta, g, pre, n, bs, acc, eps, scale, scal, vsq_slot, eps_slot, theta_p):
    _call("misc", 0.0, "byte", "fb_fd_perturb_ex", theta.data_ptr(), g.data_ptr(), L.ptr(pre), n, bs, acc, eps, scale,
          scal.data_ptr(), vsq_slot, eps_slot, theta_p.data_ptr())


def fd_combine_ex(g, g_plus, g_minus, avg, n, scal, eps_slot, cf, cursor, count0, write_g, cf_slot=-1):
    _call("misc", 0.0, "byte", "fb_fd_combine_ex", g.data_ptr(), g_plus.data_ptr(), g_minus.data_ptr(), L.ptr(avg), n,
          scal.data_ptr(), eps_slot, cf, cf_slot, L.ptr(cursor), count0, int(write_g))


def mean_accumulate_clip(g, avg, n, cursor, count0, scal, norm_slot, clip, clipped_slot):
    _call("misc", 0.0, "byte", "fb_mean_accumulate_clip", g.data_ptr(), avg.data_ptr(), n, L.ptr(cursor), count0,
          scal.data_ptr(), norm_slot, clip, clipped_slot)
