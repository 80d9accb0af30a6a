"""Per-kernel parity tests on the B200: every exported kernel of the C ABI against plain torch ops (fp64) on the same
seeded inputs, for launches of several microbatch groups.  Tolerances are stated per test: the split (hi+lo bf16)
tensor-core path carries ~16 mantissa bits per operand, so products are exact to ~2^-16 relative and results are checked
at 3e-5 of the output scale; the plain-bf16 path is checked against torch run on bf16-rounded operands (products then
exact in fp32).  Group independence is checked bit for bit: group g of a launch of G groups equals a launch of that
group alone."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from fullbatchtraining_b200 import lib as L  # noqa: E402
from fullbatchtraining_b200 import ops  # noqa: E402

DEV = torch.device("cuda")
EPS = 1e-5


def split(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


def nhwc(x):  # NCHW -> NHWC contiguous
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def rel_err(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


class ConvCase:
    """A Conv2dPlan over G groups of mb images with one shared weight and G per-group weights."""

    def __init__(self, mb, G, h, w, cin, cout, k, stride, use_split, seed=0, dx=True, x=None, weights=None, gy=None,
                 allow_pair=True):
        g = torch.Generator(device="cuda").manual_seed(seed)
        n = G * mb
        self.mb, self.G, self.cin, self.cout, self.k, self.stride, self.h, self.w = mb, G, cin, cout, k, stride, h, w
        self.x = torch.randn(n, cin, h, w, device=DEV, generator=g) if x is None else x
        sc = (2.0 / (cout * k * k)) ** 0.5
        # weights[0]: shared; weights[1 + g]: group g's own (the perturbed point of the FD pass)
        self.weights = [torch.randn(cout, cin, k, k, device=DEV, generator=g) * sc for _ in range(G + 1)] \
            if weights is None else weights
        ho, wo = h // stride, w // stride
        self.gy = torch.randn(n, cout, ho, wo, device=DEV, generator=g) * 1e-3 if gy is None else gy
        x_hi, x_lo = split(nhwc(self.x))
        if not use_split:
            x_lo = None
        taps = k * k
        bf = dict(device=DEV, dtype=torch.bfloat16)
        wsets = []
        for rows in (1, G):
            wsets.append((torch.zeros(rows * cout, taps * cin, **bf),
                          torch.zeros(rows * cout, taps * cin, **bf) if use_split else None,
                          torch.zeros(rows * cin, taps * cout, **bf),
                          torch.zeros(rows * cin, taps * cout, **bf) if use_split else None))
        ops.weight_prep(self.weights[0], cout, cin, taps, *wsets[0])
        for gi in range(G):
            f_hi, f_lo, d_hi, d_lo = wsets[1]
            ops.weight_prep(self.weights[1 + gi], cout, cin, taps, f_hi[gi * cout:(gi + 1) * cout],
                            f_lo[gi * cout:(gi + 1) * cout] if use_split else None, d_hi[gi * cin:(gi + 1) * cin],
                            d_lo[gi * cin:(gi + 1) * cin] if use_split else None)
        self.y = torch.full((n, ho, wo, cout), float("nan"), device=DEV)
        self.dy = nhwc(self.gy).to(torch.bfloat16)
        self.dx = torch.full((n, h, w, cin), float("nan"), device=DEV) if dx else None
        self.mean = torch.zeros(G, cout, device=DEV)
        self.rstd = torch.zeros(G, cout, device=DEV)
        self.bn_batch = torch.zeros(G, 2, cout, device=DEV)
        self.plan = ops.Conv2dPlan(mb, G, h, w, cin, cout, k, stride, x_hi, x_lo, self.y, self.dy, self.dx, wsets, 0,
                                   split=use_split, bn=(self.mean, self.rstd, EPS), allow_pair=allow_pair)
        self.use_split = use_split
        self.gstride = (cout * taps * cin + 63) // 64 * 64
        self.gbuf = torch.full((G, self.gstride), float("nan"), device=DEV)
        self.reduce = None
        if not self.plan.direct:
            self.partial = torch.full((self.plan.partial_elems(),), float("nan"), device=DEV)
            self.reduce = ops.ReduceTable([self.plan.bind_partial(self.partial)], DEV)

    def wgrad(self, ng):
        self.plan.wgrad(ng, self.gbuf, self.gstride)
        if self.reduce is not None:
            self.reduce(ng, self.gbuf, self.gstride)

    def operands(self, weight):
        """The values the kernel really multiplies: split -> (hi+lo) of each operand (lo*lo term below tolerance)."""
        if self.use_split:
            return self.x.double(), weight.double()
        return self.x.to(torch.bfloat16).double(), weight.to(torch.bfloat16).double()


CONV_CASES = [
    # mb, G, h, w, cin, cout, k, stride
    (4, 2, 32, 32, 64, 64, 3, 1),
    (2, 3, 32, 32, 64, 128, 3, 2),
    (4, 2, 16, 16, 128, 128, 3, 1),
    (8, 2, 8, 8, 128, 256, 1, 1),
    (4, 3, 16, 16, 128, 256, 3, 2),
    (16, 2, 4, 4, 256, 512, 3, 1),
    (8, 1, 4, 4, 512, 512, 3, 1),
    (8, 2, 8, 8, 256, 512, 3, 2),
    (16, 2, 8, 8, 64, 256, 1, 1),     # bottleneck 1x1 expansion, N tile 256
    (2, 2, 32, 32, 128, 64, 3, 1),    # 2 channel blocks
    (3, 3, 16, 16, 64, 256, 3, 1),    # 2 N tiles, odd image count
    (8, 2, 8, 8, 256, 256, 3, 1),
    (8, 4, 4, 4, 128, 128, 3, 1),
]
IDS = ["x".join(map(str, c)) for c in CONV_CASES]


@pytest.fixture(params=["halo", "pertap"], autouse=True)
def wgrad_path(request, monkeypatch):
    """wgrad fetches haloed X boxes on 32x32 / 16x16 maps by default; the per-tap variant must stay correct."""
    if request.param == "pertap":
        if "wgrad" not in request.node.name:
            pytest.skip("only wgrad tests depend on the X-fetch variant")
        monkeypatch.setenv("FB_WGRAD_HALO", "0")
    else:
        monkeypatch.setenv("FB_WGRAD_HALO", "1")
    return request.param


@pytest.mark.parametrize("use_split", [True, False], ids=["split", "bf16"])
@pytest.mark.parametrize("wset", [0, 1], ids=["shared", "pergroup"])
@pytest.mark.parametrize("case", CONV_CASES, ids=IDS)
def test_conv_forward_with_statistics(case, wset, use_split):
    mb, G, h, w, cin, cout, k, stride = case
    c = ConvCase(mb, G, h, w, cin, cout, k, stride, use_split)
    for _ in range(2):  # twice: the tickets must reset themselves
        c.mean.fill_(float("nan"))
        c.plan.forward(G, wset, c.bn_batch.data_ptr())
    torch.cuda.synchronize()
    assert torch.isfinite(c.y).all()
    for g in range(G):
        xr, wr = c.operands(c.weights[0] if wset == 0 else c.weights[1 + g])
        ref = nhwc(F.conv2d(xr[g * mb:(g + 1) * mb], wr, None, stride, (k - 1) // 2))
        got = c.y[g * mb:(g + 1) * mb]
        assert rel_err(got, ref) < 3e-5
        # BatchNorm statistics of the group's output (biased variance for rstd, unbiased for the running-stat EMA)
        yd = got.double().reshape(-1, cout)
        m, v = yd.mean(0), yd.var(0, unbiased=False)
        scale = float(yd.abs().max())  # fp32 column sums: error relative to the summands, not to the (near-zero) mean
        assert float((c.mean[g].double() - m).abs().max()) < 2e-6 * scale
        assert rel_err(c.rstd[g], 1 / (v + EPS).sqrt()) < 1e-5
        assert torch.equal(c.bn_batch[g, 0], c.mean[g])
        assert rel_err(c.bn_batch[g, 1], yd.var(0, unbiased=True)) < 1e-5


@pytest.mark.parametrize("use_split", [True, False], ids=["split", "bf16"])
@pytest.mark.parametrize("wset", [0, 1], ids=["shared", "pergroup"])
@pytest.mark.parametrize("case", CONV_CASES, ids=IDS)
def test_conv_dgrad(case, wset, use_split):
    mb, G, h, w, cin, cout, k, stride = case
    c = ConvCase(mb, G, h, w, cin, cout, k, stride, use_split)
    c.plan.dgrad(G, wset)
    torch.cuda.synchronize()
    assert torch.isfinite(c.dx).all()
    gy = nchw(c.dy.double())  # the kernel consumes the bf16-rounded output gradient
    for g in range(G):
        _, wr = c.operands(c.weights[0] if wset == 0 else c.weights[1 + g])
        ref = nhwc(torch.nn.grad.conv2d_input((mb, cin, h, w), wr, gy[g * mb:(g + 1) * mb], stride, (k - 1) // 2))
        assert rel_err(c.dx[g * mb:(g + 1) * mb], ref) < 3e-5


@pytest.mark.parametrize("use_split", [True, False], ids=["split", "bf16"])
@pytest.mark.parametrize("case", CONV_CASES, ids=IDS)
def test_conv_wgrad(case, use_split, wgrad_path):
    mb, G, h, w, cin, cout, k, stride = case
    c = ConvCase(mb, G, h, w, cin, cout, k, stride, use_split, dx=False)
    c.wgrad(G)
    torch.cuda.synchronize()
    xr, _ = c.operands(c.weights[0])
    gy = nchw(c.dy.double())
    for g in range(G):
        ref = torch.nn.grad.conv2d_weight(xr[g * mb:(g + 1) * mb], (cout, cin, k, k), gy[g * mb:(g + 1) * mb], stride,
                                          (k - 1) // 2)
        got = c.gbuf[g, :cout * k * k * cin].view(cout, k * k, cin)  # native layout [co][tap][ci]
        assert torch.isfinite(got).all()
        assert rel_err(got, ref.permute(0, 2, 3, 1).reshape(cout, k * k, cin)) < 3e-5


@pytest.mark.parametrize("case", [(4, 3, 32, 32, 64, 64, 3, 1), (4, 3, 16, 16, 128, 256, 3, 2),
                                  (8, 3, 4, 4, 256, 512, 3, 1), (8, 2, 8, 8, 256, 128, 1, 1)],
                         ids=lambda c: "x".join(map(str, c)))
def test_groups_are_independent_bit_for_bit(case):
    """Group g of a launch of G groups == a launch of that group alone (forward, statistics, dgrad, wgrad)."""
    mb, G, h, w, cin, cout, k, stride = case
    c = ConvCase(mb, G, h, w, cin, cout, k, stride, True)
    c.plan.forward(G, 1, c.bn_batch.data_ptr())
    c.plan.dgrad(G, 1)
    c.wgrad(G)
    for g in range(G):
        sl = slice(g * mb, (g + 1) * mb)
        one = ConvCase(mb, 1, h, w, cin, cout, k, stride, True, x=c.x[sl].contiguous(),
                       weights=[c.weights[1 + g], c.weights[1 + g]], gy=c.gy[sl].contiguous())
        one.plan.forward(1, 0, one.bn_batch.data_ptr())
        one.plan.dgrad(1, 0)
        one.wgrad(1)
        torch.cuda.synchronize()
        assert torch.equal(one.y, c.y[sl])
        assert torch.equal(one.mean[0], c.mean[g]) and torch.equal(one.rstd[0], c.rstd[g])
        assert torch.equal(one.bn_batch[0], c.bn_batch[g])
        assert torch.equal(one.dx, c.dx[sl])
        nw = cout * k * k * cin
        assert torch.equal(one.gbuf[0, :nw], c.gbuf[g, :nw])
    # a shorter launch (ng < G) touches only its groups
    c.y.fill_(float("nan"))
    c.plan.forward(1, 1, c.bn_batch.data_ptr())
    torch.cuda.synchronize()
    assert torch.isfinite(c.y[:mb]).all() and torch.isnan(c.y[mb:]).all()


@pytest.mark.parametrize("case", [(8, 3, 32, 32, 64, 64, 3, 1), (8, 2, 16, 16, 128, 128, 3, 1), (16, 2, 8, 8, 256, 512, 3, 2),
                                  (16, 3, 4, 4, 512, 512, 3, 1), (8, 2, 16, 16, 64, 256, 1, 1), (4, 2, 32, 32, 64, 128, 3, 2)],
                         ids=lambda c: "x".join(map(str, c)))
@pytest.mark.parametrize("use_split", [True, False], ids=["split", "bf16"])
def test_cta_pairs_match_single_cta_tiles(case, use_split):
    """tcgen05 cta_group::2 (M = 256 tiles over two SMs, half a weight tile per CTA) against the single-CTA launch of the
    same problem: forward with statistics and dgrad, shared and per-group weights.  Plain-bf16 operands: bit for bit.
    Split operands: the single-CTA kernel accumulates the hi and lo weight planes in two accumulators that the epilogue
    adds, a pair accumulates all products in one -- the same products in another fp32 order (2e-5 of the output scale)."""
    mb, G, h, w, cin, cout, k, stride = case
    a = ConvCase(mb, G, h, w, cin, cout, k, stride, use_split, allow_pair="force")
    b = ConvCase(mb, G, h, w, cin, cout, k, stride, use_split, allow_pair=False)
    assert a.plan.pair_fwd and not b.plan.pair_fwd, "the case is meant to exercise CTA pairs"
    for wset in (0, 1):
        for c in (a, b):
            c.y.fill_(float("nan"))
            c.dx.fill_(float("nan"))
            c.plan.forward(G, wset, c.bn_batch.data_ptr())
            c.plan.dgrad(G, wset)
        torch.cuda.synchronize()
        assert torch.isfinite(a.y).all() and torch.isfinite(a.dx).all()
        if use_split:
            assert rel_err(a.y, b.y) < 2e-5 and rel_err(a.dx, b.dx) < 2e-5  # fp32 sums of up to 4,608 x 3 products
            assert float((a.mean - b.mean).abs().max()) < 1e-6 * float(b.y.abs().max()) and rel_err(a.rstd, b.rstd) < 1e-5
        else:
            assert torch.equal(a.y, b.y) and torch.equal(a.dx, b.dx)
            assert torch.equal(a.mean, b.mean) and torch.equal(a.rstd, b.rstd) and torch.equal(a.bn_batch, b.bn_batch)


@pytest.mark.parametrize("case", [(8, 3, 32, 32, 64, 64, 3, 1), (8, 2, 16, 16, 128, 128, 3, 1), (16, 2, 8, 8, 256, 512, 3, 2),
                                  (16, 3, 4, 4, 512, 512, 3, 1), (8, 2, 16, 16, 64, 256, 1, 1), (4, 2, 32, 32, 64, 128, 3, 2)],
                         ids=lambda c: "x".join(map(str, c)))
@pytest.mark.parametrize("use_split", [True, False], ids=["split", "bf16"])
def test_weight_multicast_is_bit_identical_to_independent_ctas(case, use_split):
    """Clusters of two CTAs sharing every weight tile by TMA multicast (cta_pair = 2) issue exactly the instructions of
    two independent CTAs: outputs, fused BatchNorm statistics and dgrad are bit for bit those of the plain launch."""
    mb, G, h, w, cin, cout, k, stride = case
    a = ConvCase(mb, G, h, w, cin, cout, k, stride, use_split, allow_pair="mcast")
    b = ConvCase(mb, G, h, w, cin, cout, k, stride, use_split, allow_pair=False)
    assert a.plan.pair_fwd == 2 and not b.plan.pair_fwd, "the case is meant to exercise the multicast mode"
    for wset in (0, 1):
        for c in (a, b):
            c.y.fill_(float("nan"))
            c.dx.fill_(float("nan"))
            c.plan.forward(G, wset, c.bn_batch.data_ptr())
            c.plan.dgrad(G, wset)
        torch.cuda.synchronize()
        assert torch.isfinite(a.y).all() and torch.isfinite(a.dx).all()
        assert torch.equal(a.y, b.y) and torch.equal(a.dx, b.dx)
        assert torch.equal(a.mean, b.mean) and torch.equal(a.rstd, b.rstd) and torch.equal(a.bn_batch, b.bn_batch)


@pytest.mark.parametrize("case", [(8, 3, 32, 32, 64, 64, 3, 1), (8, 2, 16, 16, 128, 128, 3, 1), (4, 2, 32, 32, 128, 64, 3, 1),
                                  (8, 1, 16, 16, 64, 128, 3, 1), (3, 2, 32, 32, 64, 64, 3, 1)],
                         ids=lambda c: "x".join(map(str, c)))
@pytest.mark.parametrize("use_split", [True, False], ids=["split", "bf16"])
def test_haloed_boxes_match_per_tap_boxes(case, use_split):
    """3x3 / stride-1 forward and dgrad with ONE haloed A box per filter column (fb_conv_gemm_args.halo) against the
    launch that fetches a box per tap: the same products summed column-major instead of row-major over the taps (fp32
    order: 2e-5 of the output scale), fused BatchNorm statistics included; per-group weights; image borders."""
    mb, G, h, w, cin, cout, k, stride = case
    a = ConvCase(mb, G, h, w, cin, cout, k, stride, use_split, allow_pair="halo")
    b = ConvCase(mb, G, h, w, cin, cout, k, stride, use_split, allow_pair="nohalo")
    assert a.plan.halo_fwd and a.plan.halo_dgrad and not b.plan.halo_fwd and not b.plan.halo_dgrad
    for wset in (0, 1):
        for c in (a, b):
            c.y.fill_(float("nan"))
            c.dx.fill_(float("nan"))
            c.plan.forward(G, wset, c.bn_batch.data_ptr())
            c.plan.dgrad(G, wset)
        torch.cuda.synchronize()
        assert torch.isfinite(a.y).all() and torch.isfinite(a.dx).all()
        assert rel_err(a.y, b.y) < 2e-5 and rel_err(a.dx, b.dx) < 2e-5
        assert float((a.mean - b.mean).abs().max()) < 1e-6 * float(b.y.abs().max()) and rel_err(a.rstd, b.rstd) < 1e-5
    # twice the same launch: bit-identical
    y0 = a.y.clone()
    a.plan.forward(G, 1, a.bn_batch.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(y0, a.y)


def test_weight_prep_layouts():
    g = torch.Generator(device="cuda").manual_seed(3)
    cout, cin = 128, 64
    w = torch.randn(cout, cin, 3, 3, device=DEV, generator=g)
    wf_hi = torch.zeros(cout, 9 * cin, device=DEV, dtype=torch.bfloat16)
    wf_lo = torch.zeros_like(wf_hi)
    wd_hi = torch.zeros(cin, 9 * cout, device=DEV, dtype=torch.bfloat16)
    wd_lo = torch.zeros_like(wd_hi)
    ops.weight_prep(w, cout, cin, 9, wf_hi, wf_lo, wd_hi, wd_lo)
    hi, lo = split(w)
    exp_f = hi.permute(0, 2, 3, 1).reshape(cout, 9 * cin)
    assert torch.equal(wf_hi, exp_f)
    assert torch.equal(wf_lo, lo.permute(0, 2, 3, 1).reshape(cout, 9 * cin))
    assert torch.equal(wd_hi, hi.permute(1, 2, 3, 0).reshape(cin, 9 * cout))
    assert torch.equal(wd_lo, lo.permute(1, 2, 3, 0).reshape(cin, 9 * cout))
    # hi + lo reproduces fp32 to ~2^-17
    assert float(((hi.float() + lo.float()) - w).abs().max() / w.abs().max()) < 2e-5


def close_split(hi, lo, e_hi, e_lo):
    """hi + lo equal up to the rounding of the perturbed fp32 value (the kernel contracts theta + step*v into an FMA):
    1 ulp of fp32 on the value, which can move the rounding of lo by one step = the 2^-17 resolution of the pair"""
    a, b = hi.double() + lo.double(), e_hi.double() + e_lo.double()
    return float((a - b).abs().max()) <= 2.0 ** -16 * float(b.abs().max())


def test_weight_prep_multi_perturbed_points_and_relayout():
    """fb_weight_prep_multi: operands of theta + scale*eps_g*(bs*grad_g + acc*pre) for every group, gradient in native
    layout, against fb_weight_prep of the explicitly perturbed weight; fb_flat_relayout round trip."""
    gen = torch.Generator(device="cuda").manual_seed(13)
    G = 3
    shapes = [("stem", 64, 3, 9), ("c1", 64, 64, 9), ("c2", 128, 64, 1), ("c3", 128, 128, 9)]
    offs, off = {}, 0
    for name, co, ci, t in shapes:
        offs[name] = off
        off += co * ci * t
        off += 128  # some non-conv parameters in between (BatchNorm weight / bias)
    numel = off
    stride = (numel + 63) // 64 * 64
    theta = torch.randn(numel, device=DEV, generator=gen)
    grad_oihw = torch.randn(G, numel, device=DEV, generator=gen) * 1e-2
    pre_oihw = torch.randn(numel, device=DEV, generator=gen) * 1e-2
    table = torch.tensor([[offs["c1"], 64, 64, 9, 0], [offs["c3"], 128, 128, 9, 64]], dtype=torch.int64, device=DEV)
    grad_n = torch.zeros(G, stride, device=DEV)
    pre_n = torch.zeros(numel, device=DEV)
    for g in range(G):
        ops.flat_relayout(grad_oihw[g], grad_n[g, :numel], numel, table, 2, 192, True)
    ops.flat_relayout(pre_oihw, pre_n, numel, table, 2, 192, True)
    back = torch.zeros(numel, device=DEV)
    ops.flat_relayout(grad_n[1, :numel], back, numel, table, 2, 192, False)
    assert torch.equal(back, grad_oihw[1])
    c1 = grad_oihw[0, offs["c1"]:offs["c1"] + 64 * 64 * 9].view(64, 64, 9)
    assert torch.equal(grad_n[0, offs["c1"]:offs["c1"] + 64 * 64 * 9].view(64, 9, 64), c1.permute(0, 2, 1))
    scal = torch.zeros(64, device=DEV)
    eps_base = 16
    scal[eps_base:eps_base + G] = torch.tensor([0.3, 0.7, 1.1], device=DEV)
    bs, acc, scale = 0.5, 0.25, -0.5
    bf = dict(device=DEV, dtype=torch.bfloat16)
    entries, bufs = [], {}
    for name, co, ci, t in shapes:
        cin_k = 64 if name == "stem" else ci
        kk = 1 if name == "stem" else t
        wf = [torch.zeros(G * co, kk * cin_k, **bf) for _ in range(2)]
        wd = [torch.zeros(G * ci, t * co, **bf) for _ in range(2)] if name != "stem" else [None, None]
        bufs[name] = (wf, wd)
        entries.append((offs[name], co, ci, t, wf[0], wf[1], wd[0], wd[1]))
    tab = ops.WeightPrepTable(entries, DEV, per_group=True)
    tab(theta, ng=G, grad=grad_n, gstride=stride, pre=pre_n, bs=bs, acc=acc, scale=scale, scal=scal, eps_base=eps_base)
    torch.cuda.synchronize()
    for name, co, ci, t in shapes:
        wf, wd = bufs[name]
        for g in range(G):
            sl = slice(offs[name], offs[name] + co * ci * t)
            step = scale * float(scal[eps_base + g])
            wp = (theta[sl] + step * (bs * grad_oihw[g, sl] + acc * pre_oihw[sl])).view(co, ci, 3 if t == 9 else 1,
                                                                                         3 if t == 9 else 1)
            if name == "stem":
                e_hi = torch.zeros(co, 64, **bf)
                e_lo = torch.zeros(co, 64, **bf)
                ops.weight_prep(wp.contiguous(), co, ci, t, e_hi, e_lo)
                assert close_split(wf[0][g * co:(g + 1) * co], wf[1][g * co:(g + 1) * co], e_hi, e_lo)
                continue
            e = [torch.zeros(co, t * ci, **bf), torch.zeros(co, t * ci, **bf), torch.zeros(ci, t * co, **bf),
                 torch.zeros(ci, t * co, **bf)]
            ops.weight_prep(wp.contiguous(), co, ci, t, *e)
            assert close_split(wf[0][g * co:(g + 1) * co], wf[1][g * co:(g + 1) * co], e[0], e[1])
            assert close_split(wd[0][g * ci:(g + 1) * ci], wd[1][g * ci:(g + 1) * ci], e[2], e[3])
    # plain theta (no gradient): one shared operand set
    entries0 = []
    for name, co, ci, t in shapes[1:]:
        wf = [torch.zeros(co, t * ci, **bf) for _ in range(2)]
        wd = [torch.zeros(ci, t * co, **bf) for _ in range(2)]
        entries0.append((offs[name], co, ci, t, wf[0], wf[1], wd[0], wd[1]))
    ops.WeightPrepTable(entries0, DEV, per_group=False)(theta)
    for (name, co, ci, t), ent in zip(shapes[1:], entries0):
        e = [torch.zeros_like(ent[4]), torch.zeros_like(ent[5]), torch.zeros_like(ent[6]), torch.zeros_like(ent[7])]
        kk = 3 if t == 9 else 1
        ops.weight_prep(theta[offs[name]:offs[name] + co * ci * t].view(co, ci, kk, kk), co, ci, t, *e)
        for a, b in zip(ent[4:], e):
            assert torch.equal(a, b)


def test_stem_im2col_and_stem_conv():
    g = torch.Generator(device="cuda").manual_seed(5)
    mb, G = 4, 2
    n = mb * G
    data = torch.randn(40, 3, 32, 32, device=DEV, generator=g)
    labels = torch.randint(0, 10, (40,), device=DEV, generator=g)
    perm = torch.randperm(40, device=DEV, generator=g)
    cursor = torch.tensor([3], device=DEV, dtype=torch.int32)
    p_hi = torch.empty(n * 1024, 64, device=DEV, dtype=torch.bfloat16)
    p_lo = torch.empty_like(p_hi)
    lab = torch.empty(n, device=DEV, dtype=torch.int64)
    ops.stem_im2col(data, labels, perm, cursor, 8, mb, n, p_hi, p_lo, lab)  # samples perm[8 + 3*4 : 8 + 3*4 + 8]
    idx = perm[20:28]
    x = data[idx]
    assert torch.equal(lab, labels[idx])
    patches = F.unfold(x, 3, padding=1).transpose(1, 2).reshape(n * 1024, 27)  # column = ci*9 + kh*3 + kw
    got = p_hi.float() + p_lo.float()
    assert float((got[:, :27] - patches).abs().max()) < 2e-5 * float(patches.abs().max())
    assert float(got[:, 27:].abs().max()) == 0.0
    # stem conv = 1x1 GEMM over the patches with the OIHW weight as B (27 of 64 columns used)
    w = torch.randn(64, 3, 3, 3, device=DEV, generator=g) * 0.1
    bf = dict(device=DEV, dtype=torch.bfloat16)
    wsets = [(torch.zeros(64, 64, **bf), torch.zeros(64, 64, **bf), None, None),
             (torch.zeros(G * 64, 64, **bf), torch.zeros(G * 64, 64, **bf), None, None)]
    ops.weight_prep(w, 64, 3, 9, wsets[0][0], wsets[0][1])
    y = torch.empty(n, 32, 32, 64, device=DEV)
    gy = torch.randn(n, 64, 32, 32, device=DEV, generator=g) * 1e-3
    dy = nhwc(gy).to(torch.bfloat16)
    w_off = 192
    plan = ops.Conv2dPlan(mb, G, 32, 32, 64, 64, 1, 1, p_hi.view(n, 32, 32, 64), p_lo.view(n, 32, 32, 64), y, dy, None,
                          wsets, w_off, alg_k=27, grad_cols=27)
    plan.forward(G, 0)
    ref = nhwc(F.conv2d(x.double(), w.double(), None, 1, 1))
    assert rel_err(y, ref) < 3e-5
    assert not plan.direct
    partial = torch.full((plan.partial_elems(),), float("nan"), device=DEV)
    red = ops.ReduceTable([plan.bind_partial(partial)], DEV)
    stride = 4096
    gbuf = torch.full((G, stride), float("nan"), device=DEV)
    plan.wgrad(G, gbuf, stride)
    red(G, gbuf, stride)
    torch.cuda.synchronize()
    for gi in range(G):
        sl = slice(gi * mb, (gi + 1) * mb)
        refw = torch.nn.grad.conv2d_weight(x[sl].double(), (64, 3, 3, 3), nchw(dy[sl].double()), 1, 1)
        assert rel_err(gbuf[gi, w_off:w_off + 64 * 27].view(64, 3, 3, 3), refw) < 3e-5
        assert torch.isnan(gbuf[gi, :w_off]).all() and torch.isnan(gbuf[gi, w_off + 64 * 27:]).all()


@pytest.mark.parametrize("P,Cc", [(4096, 64), (131072, 64), (2048, 512), (512, 2048), (1000, 128)])
def test_bn_stats(P, Cc):
    g = torch.Generator(device="cuda").manual_seed(P + Cc)
    y = torch.randn(P, Cc, device=DEV, generator=g) * 2 + 0.5
    ws = torch.zeros(2 * Cc * 1024, device=DEV)
    mean, rstd = torch.empty(Cc, device=DEV), torch.empty(Cc, device=DEV)
    rm, rv = torch.zeros(Cc, device=DEV), torch.ones(Cc, device=DEV)
    ops.bn_stats(y, P, Cc, ws, mean, rstd, rm, rv)
    yd = y.double()
    m, v = yd.mean(0), yd.var(0, unbiased=False)
    assert rel_err(mean, m) < 1e-6
    assert rel_err(rstd, 1 / (v + 1e-5).sqrt()) < 1e-6
    assert rel_err(rm, 0.1 * m) < 1e-6
    assert rel_err(rv, 0.9 + 0.1 * yd.var(0, unbiased=True)) < 1e-6


def group_stats(y, ng, P, Cc):
    yd = y.double().view(ng, P, Cc)
    mean = yd.mean(1)
    rstd = 1 / (yd.var(1, unbiased=False) + EPS).sqrt()
    return mean.float().contiguous(), rstd.float().contiguous()


@pytest.mark.parametrize("variant", ["plain", "residual", "dual", "norelu"])
@pytest.mark.parametrize("shape", [(3, 4, 8, 8, 128), (2, 2, 32, 32, 64), (2, 8, 4, 4, 512), (1, 5, 2, 2, 2048)],
                         ids=lambda s: "x".join(map(str, s)))
def test_bn_apply_and_backward(variant, shape):
    """fb_bn_apply / fb_bn_bwd over ng groups with per-group statistics AND per-group (perturbed) affine parameters,
    against torch.nn.functional.batch_norm + autograd in fp64, group by group."""
    gen = torch.Generator(device="cuda").manual_seed(11)
    ng, n, h, w, Cc = shape
    P = n * h * w
    pstride = 4 * Cc + 64  # gamma | beta | gamma2 | beta2 of a group, like rows of theta_p
    params = torch.randn(ng, pstride, device=DEV, generator=gen) * 0.1
    params[:, :Cc] += 1.0
    params[:, 2 * Cc:3 * Cc] += 1.0
    y = torch.randn(ng * P, Cc, device=DEV, generator=gen) * 1.5 + 0.3
    y2 = torch.randn(ng * P, Cc, device=DEV, generator=gen)
    res_hi, res_lo = split(torch.randn(ng * P, Cc, device=DEV, generator=gen))
    mean, rstd = group_stats(y, ng, P, Cc)
    mean2, rstd2 = group_stats(y2, ng, P, Cc)
    out_hi = torch.empty(ng * P, Cc, device=DEV, dtype=torch.bfloat16)
    out_lo = torch.empty_like(out_hi)
    relu = variant != "norelu"
    base = params.data_ptr()
    bits = torch.zeros(ng * P * Cc // 8, device=DEV, dtype=torch.uint8)
    ops.bn_apply(y, mean, rstd, base, base + 4 * Cc, P, Cc, out_hi, out_lo, relu=relu,
                 second=(y2, mean2, rstd2, base + 8 * Cc, base + 12 * Cc) if variant == "dual" else None,
                 res=(res_hi, res_lo) if variant == "residual" else None, ng=ng, param_gstride=pstride, mask_out=bits)
    # the ReLU mask as a bit plane: byte e/8, bit e%8
    exp_bits = ((out_hi.float() + out_lo.float()) > 0).view(-1, 8).to(torch.uint8)
    exp_bits = (exp_bits << torch.arange(8, device=DEV, dtype=torch.uint8)).sum(1).to(torch.uint8)
    assert torch.equal(bits, exp_bits)
    dA = torch.randn(ng * P, Cc, device=DEV, generator=gen)
    dA2 = torch.randn(ng * P, Cc, device=DEV, generator=gen) if variant in ("residual", "plain") else None
    gstride = 2 * Cc + 64
    grads = torch.full((ng, gstride), float("nan"), device=DEV)
    dy = torch.empty(ng * P, Cc, device=DEV, dtype=torch.bfloat16)
    dz = torch.empty(ng * P, Cc, device=DEV)
    ws = torch.zeros(ops.bn_bwd_ws_floats(P, Cc, ng), device=DEV)
    for _ in range(2):  # twice: the tickets must reset themselves
        ops.bn_bwd(dA, out_hi if relu else None, y, mean, rstd, base, P, Cc, ws, grads.data_ptr(),
                   grads.data_ptr() + 4 * Cc, dy, dz_out=dz, dA2=dA2, ng=ng, param_gstride=pstride, grad_gstride=gstride)
    if relu:  # the bit plane instead of the bf16 plane: identical results
        grads_b = torch.full((ng, gstride), float("nan"), device=DEV)
        dy_b, dz_b = torch.empty_like(dy), torch.empty_like(dz)
        ops.bn_bwd(dA, None, y, mean, rstd, base, P, Cc, ws, grads_b.data_ptr(), grads_b.data_ptr() + 4 * Cc, dy_b,
                   dz_out=dz_b, dA2=dA2, ng=ng, param_gstride=pstride, grad_gstride=gstride, mask_bits=bits)
        assert torch.equal(dy_b, dy) and torch.equal(dz_b, dz) and torch.equal(grads_b[:, :2 * Cc], grads[:, :2 * Cc])
    # dz written over dA in place (projection-shortcut blocks), then consumed premasked by a second BatchNorm backward
    # without addend or mask: the same dY / dgamma / dbeta bits as the masked two-addend call
    dA_ip, grads_ip, dy_ip = dA.clone(), torch.full((ng, gstride), float("nan"), device=DEV), torch.empty_like(dy)
    ops.bn_bwd(dA_ip, out_hi if relu else None, y, mean, rstd, base, P, Cc, ws, grads_ip.data_ptr(),
               grads_ip.data_ptr() + 4 * Cc, dy_ip, dz_out=dA_ip, dA2=dA2, ng=ng, param_gstride=pstride, grad_gstride=gstride)
    assert torch.equal(dA_ip, dz) and torch.equal(dy_ip, dy) and torch.equal(grads_ip[:, :2 * Cc], grads[:, :2 * Cc])
    grads_pm, dy_pm = torch.full((ng, gstride), float("nan"), device=DEV), torch.empty_like(dy)
    ops.bn_bwd(dA_ip, None, y, mean, rstd, base, P, Cc, ws, grads_pm.data_ptr(), grads_pm.data_ptr() + 4 * Cc, dy_pm,
               ng=ng, param_gstride=pstride, grad_gstride=gstride)
    assert torch.equal(dy_pm, dy) and torch.equal(grads_pm[:, :2 * Cc], grads[:, :2 * Cc])
    torch.cuda.synchronize()
    for g in range(ng):
        sl = slice(g * P, (g + 1) * P)
        yd = y[sl].double().requires_grad_(True)
        y2d = y2[sl].double()
        resd = (res_hi[sl].double() + res_lo[sl].double()).requires_grad_(True)
        ga = params[g, :Cc].double().requires_grad_(True)
        be = params[g, Cc:2 * Cc].double().requires_grad_(True)

        def bn(t, gamma, beta):
            t4 = t.view(n, h, w, Cc).permute(0, 3, 1, 2)
            o = F.batch_norm(t4, None, None, gamma, beta, True, 0.1, EPS)
            return o.permute(0, 2, 3, 1).reshape(P, Cc)

        ref = bn(yd, ga, be)
        if variant == "dual":
            ref = ref + bn(y2d, params[g, 2 * Cc:3 * Cc].double(), params[g, 3 * Cc:4 * Cc].double())
        if variant == "residual":
            ref = ref + resd
        if relu:
            ref = F.relu(ref)
        got = out_hi[sl].double() + out_lo[sl].double()
        assert rel_err(got, ref) < 2e-5
        up = dA[sl].double() + (dA2[sl].double() if dA2 is not None else 0.0)
        gy_ref, gga, gbe = torch.autograd.grad(ref, (yd, ga, be), up, retain_graph=True)
        # dy is stored in bf16 (8 mantissa bits): 2^-8 relative to the output scale
        assert rel_err(dy[sl], gy_ref) < 5e-3
        mask = (ref > 0).double() if relu else torch.ones_like(ref)
        assert rel_err(dz[sl], up * mask) < 1e-6
        assert rel_err(grads[g, :Cc], gga) < 2e-5
        assert rel_err(grads[g, Cc:2 * Cc], gbe) < 2e-5
        if variant == "residual":
            gres, = torch.autograd.grad(ref, resd, up)
            assert rel_err(dz[sl], gres) < 1e-6
    # group independence, bit for bit: group 1 alone
    if ng > 1:
        sl = slice(P, 2 * P)
        g1 = torch.full((1, gstride), float("nan"), device=DEV)
        dy1 = torch.empty(P, Cc, device=DEV, dtype=torch.bfloat16)
        ws1 = torch.zeros(ops.bn_bwd_ws_floats(P, Cc, 1), device=DEV)
        ops.bn_bwd(dA[sl], out_hi[sl] if relu else None, y[sl], mean[1:2].contiguous(), rstd[1:2].contiguous(),
                   base + 4 * pstride, P, Cc, ws1, g1.data_ptr(), g1.data_ptr() + 4 * Cc, dy1,
                   dA2=dA2[sl] if dA2 is not None else None, ng=1)
        assert torch.equal(dy1, dy[sl]) and torch.equal(g1[0, :2 * Cc], grads[1, :2 * Cc])


def test_bn_ema_multi_follows_the_reference_order():
    gen = torch.Generator(device="cuda").manual_seed(17)
    G, passes, mom = 4, 2, 0.1
    layers = [64, 128, 64]
    entries, refs = [], []
    for Cc in layers:
        rm, rv = torch.randn(Cc, device=DEV, generator=gen), torch.rand(Cc, device=DEV, generator=gen) + 0.5
        batch = torch.randn(3, G, 2, Cc, device=DEV, generator=gen)
        refs.append((rm.clone(), rv.clone(), batch))
        entries.append((rm, rv, batch, batch.stride(0), Cc))
    tab = ops.BnEmaTable(entries, DEV)
    tab(passes, 3, mom)  # three of the four groups
    torch.cuda.synchronize()
    for (rm, rv, _, _, _), (rm0, rv0, batch) in zip(entries, refs):
        for g in range(3):
            for p in range(passes):  # microbatch-major: pass 1 of k, pass 2 of k, pass 1 of k+1, ... (training.py:148-173)
                rm0 = (1 - mom) * rm0 + mom * batch[p, g, 0]
                rv0 = (1 - mom) * rv0 + mom * batch[p, g, 1]
        # same recurrence in the same order; the kernel contracts it into FMAs -> equal to fp32 rounding
        assert torch.allclose(rm, rm0, rtol=1e-6, atol=1e-7) and torch.allclose(rv, rv0, rtol=1e-6, atol=1e-7)


def test_avgpool2():
    g = torch.Generator(device="cuda").manual_seed(2)
    n, h, w, c = 4, 16, 16, 64
    x = torch.randn(n, h, w, c, device=DEV, generator=g)
    hi, lo = split(x)
    o_hi = torch.empty(n, h // 2, w // 2, c, device=DEV, dtype=torch.bfloat16)
    o_lo = torch.empty_like(o_hi)
    ops.avgpool2_fwd(hi, lo, n, h, w, c, o_hi, o_lo)
    ref = nhwc(F.avg_pool2d(nchw(hi.double() + lo.double()), 2, 2))
    assert rel_err(o_hi.double() + o_lo.double(), ref) < 2e-5
    dP = torch.randn(n, h // 2, w // 2, c, device=DEV, generator=g)
    dX = torch.full((n, h, w, c), 1.0, device=DEV)
    ops.avgpool2_bwd(dP, n, h, w, c, dX, accumulate=True)
    refb = 1.0 + nhwc(F.interpolate(nchw(dP.double()), scale_factor=2, mode="nearest")) / 4
    assert rel_err(dX, refb) < 1e-6


@pytest.mark.parametrize("smoothing", [0.0, 0.1])
@pytest.mark.parametrize("ng", [1, 3])
def test_head(smoothing, ng):
    """pool + linear + label-smoothed CE + accuracy + backward, per group with per-group fc parameters"""
    g = torch.Generator(device="cuda").manual_seed(9)
    n, hw, c, classes = 16, 16, 512, 10
    a = torch.relu(torch.randn(ng * n, hw, c, device=DEV, generator=g))
    hi, lo = split(a)
    pstride = classes * c + 64
    params = torch.randn(ng, pstride, device=DEV, generator=g) * 0.05
    labels = torch.randint(0, classes, (ng * n,), device=DEV, generator=g)
    ws = torch.empty(ops.head_ws_floats(ng * n, c), device=DEV)
    scal = torch.zeros(64, device=DEV)
    gstride = pstride
    grads = torch.full((ng, gstride), float("nan"), device=DEV)
    dA = torch.empty(ng * n, hw, c, device=DEV)
    pb, gb = params.data_ptr(), grads.data_ptr()
    ops.head_fwd_bwd(hi, lo, n, hw, c, pb, pb + 4 * classes * c, labels, classes, smoothing, ws, scal, 16, 32,
                     gb, gb + 4 * classes * c, dA, ng=ng, param_gstride=pstride, grad_gstride=gstride)
    for k in range(ng):
        sl = slice(k * n, (k + 1) * n)
        ad = (hi[sl].double() + lo[sl].double()).requires_grad_(True)
        wd = params[k, :classes * c].view(classes, c).double().requires_grad_(True)
        bd = params[k, classes * c:classes * c + classes].double().requires_grad_(True)
        logits = F.linear(ad.mean(1), wd, bd)
        logp = F.log_softmax(logits, -1)
        wgt = torch.full_like(logits, smoothing / (classes - 1))
        wgt.scatter_(-1, labels[sl].unsqueeze(-1), 1 - smoothing)
        loss = (-wgt * logp).sum(-1).mean()
        ga, gw, gbias = torch.autograd.grad(loss, (ad, wd, bd))
        assert abs(float(scal[16 + k]) - float(loss)) < 1e-5 * abs(float(loss))
        assert float(scal[32 + k]) == float((logits.argmax(-1) == labels[sl]).sum())
        assert rel_err(dA[sl], ga) < 1e-5
        assert rel_err(grads[k, :classes * c].view(classes, c), gw) < 1e-5
        assert rel_err(grads[k, classes * c:classes * c + classes], gbias) < 1e-5


def test_flat_fd_kernels():
    gen = torch.Generator(device="cuda").manual_seed(4)
    n, G = 1_000_003, 3  # not a multiple of 4: exercises the tails
    stride = (n + 63) // 64 * 64
    theta = torch.randn(n, device=DEV, generator=gen)
    grad = torch.zeros(G, stride, device=DEV)
    grad[:, :n] = torch.randn(G, n, device=DEV, generator=gen) * 1e-3
    grad *= torch.tensor([1.0, 2.0, 3.0], device=DEV)[:, None]  # distinct norms (for the clip test below)
    pre = torch.randn(n, device=DEV, generator=gen) * 1e-3
    g2 = grad + torch.randn(G, stride, device=DEV, generator=gen) * 1e-6
    g3 = grad + torch.randn(G, stride, device=DEV, generator=gen) * 1e-6
    avg = torch.randn(n, device=DEV, generator=gen) * 1e-3
    ws = torch.empty(1024 * G, device=DEV, dtype=torch.float64)
    scal = torch.zeros(160, device=DEV)
    cursor = torch.tensor([2], device=DEV, dtype=torch.int32)
    norms = torch.zeros(8, device=DEV)
    bs, acc, eps, cf = 0.5, 0.25, 1e-2, 0.2
    N2, EP, VS = 16, 32, 48
    ops.flat_sqnorm(grad, n, ws, scal, N2, ng=G, gstride=stride, norms_out=norms, cursor=cursor, eps_mode=1, bs=bs,
                    eps=eps, eps_base=EP)
    for g in range(G):
        n2 = float(grad[g, :n].double().pow(2).sum())
        assert abs(float(scal[N2 + g]) - n2) < 1e-6 * n2
        assert abs(float(scal[EP + g]) - eps / (bs * n2 ** 0.5)) < 1e-6 * eps / (bs * n2 ** 0.5)
        assert float(norms[2 + g]) == float(scal[N2 + g])
    # |bs*g + acc*pre|^2 and eps / sqrt of it (acc_strength, modules.py:217-223)
    ops.flat_sqnorm(grad, n, ws, scal, VS, ng=G, gstride=stride, y=pre, a=bs, b=acc, eps_mode=2, eps=eps, eps_base=EP)
    for g in range(G):
        v2 = float((bs * grad[g, :n].double() + acc * pre.double()).pow(2).sum())
        assert abs(float(scal[VS + g]) - v2) < 1e-6 * v2
        assert abs(float(scal[EP + g]) - eps / v2 ** 0.5) < 1e-6 * eps / v2 ** 0.5
    # perturbation of index ranges
    ranges = torch.tensor([[64, 1000, 0], [5000, 333, 1000], [900000, 100003, 1333]], dtype=torch.int64, device=DEV)
    theta_p = torch.full((G, stride), float("nan"), device=DEV)
    ops.perturb_ranges(theta, grad, stride, pre, ranges, 3, 1000 + 333 + 100003, bs, acc, -0.5, scal, EP, theta_p, stride,
                       G)
    for g in range(G):
        step = -0.5 * float(scal[EP + g])
        for o, ln, _ in ranges.tolist():
            exp = theta[o:o + ln].double() + step * (bs * grad[g, o:o + ln].double() + acc * pre[o:o + ln].double())
            assert rel_err(theta_p[g, o:o + ln], exp) < 1e-6
        assert torch.isnan(theta_p[g, :64]).all() and torch.isnan(theta_p[g, 1064:5000]).all()
    # combine + running mean, groups in loader order (forward and central differences)
    scal[2] = cf
    for g_minus in (None, g3):
        g_in, avg_in = grad.clone(), avg.clone()
        ops.fd_combine(g_in, g2, g_minus, stride, avg_in, n, G, scal, EP, 2, cursor, True)
        ref_avg = avg.double()
        for g in range(G):
            en = float(scal[EP + g])
            base = grad[g, :n].double() if g_minus is None else g_minus[g, :n].double()
            g_reg = grad[g, :n].double() + cf * (g2[g, :n].double() - base) / en
            # (g2 - g) is formed in fp32 from nearly equal numbers: one fp32 ulp of g, amplified by cf/eps_n
            tol = float(grad.abs().max()) * 2 ** -23 * cf / en * 2
            assert float((g_in[g, :n].double() - g_reg).abs().max()) < tol
            ref_avg = ref_avg + (g_in[g, :n].double() - ref_avg) / (2 + g + 1)
        assert rel_err(avg_in, ref_avg) < 1e-6
    # plain running mean with the per-microbatch clip (training/utils.py:4-19)
    avg2, g_in = avg.clone(), grad.clone()
    norms_true = [float(grad[g, :n].double().norm()) for g in range(G)]
    clip = 0.5 * (sorted(norms_true)[0] + sorted(norms_true)[1])  # clips two of the three
    scal[3] = 0
    ops.mean_accumulate(g_in, stride, avg2, n, G, cursor, scal, N2, clip, 3)
    ref_avg = avg.double()
    for g in range(G):
        coef = clip / (norms_true[g] + 1e-6) if norms_true[g] > clip else 1.0
        ref_avg = ref_avg + (grad[g, :n].double() * coef - ref_avg) / (2 + g + 1)
    assert rel_err(avg2, ref_avg) < 1e-6
    assert float(scal[3]) == 2.0
    # end of a group launch
    scal[0], scal[1] = 1.5, 10.0
    scal[64:64 + G] = torch.tensor([0.5, 0.25, 0.125], device=DEV)
    scal[80:80 + G] = torch.tensor([3.0, 4.0, 5.0], device=DEV)
    ops.group_finish(cursor, G, scal, 0, 1, 64, 80)
    assert int(cursor) == 2 + G and float(scal[0]) == 1.5 + 0.875 and float(scal[1]) == 22.0
    # sums kept in another scalar block (shared by the lanes), cursor advanced by the lanes' stride
    totals = torch.zeros(8, device=DEV)
    ops.group_finish(cursor, G, scal, 0, 1, 64, 80, cursor_step=2 * G, totals=totals)
    assert int(cursor) == 2 + 3 * G and float(totals[0]) == 0.875 and float(totals[1]) == 12.0 and float(scal[0]) == 2.375
    x = grad[0, :n].clone()
    ops.flat_scale(x, n, 0.25)
    assert torch.equal(x, grad[0, :n] * 0.25)


def test_stem_im2col_u8_augmentation_matches_torchvision_semantics():
    """RandomCrop(32, padding=4) -> RandomHorizontalFlip -> ToTensor -> Normalize with GIVEN draws, fused into the stem
    im2col, against the same pipeline spelled out with torch ops (data_preparation.py:173-200)."""
    g = torch.Generator(device="cuda").manual_seed(8)
    N, n = 40, 8
    data = torch.randint(0, 256, (N, 32, 32, 3), device=DEV, generator=g, dtype=torch.uint8)
    labels = torch.randint(0, 10, (N,), device=DEV, generator=g)
    perm = torch.randperm(N, device=DEV, generator=g)
    aug = torch.zeros(N, 4, device=DEV, dtype=torch.int8)
    aug[:, 0:2] = torch.randint(0, 9, (N, 2), device=DEV, generator=g).to(torch.int8)
    aug[:, 2] = (torch.rand(N, device=DEV, generator=g) < 0.5).to(torch.int8)
    mean, std = [0.4914, 0.4822, 0.4465], [0.2470, 0.2435, 0.2616]
    cursor = torch.tensor([2], device=DEV, dtype=torch.int32)
    p_hi = torch.empty(n * 1024, 64, device=DEV, dtype=torch.bfloat16)
    p_lo = torch.empty_like(p_hi)
    lab = torch.empty(n, device=DEV, dtype=torch.int64)
    first = 4
    ops.stem_im2col_u8aug(data, labels, perm, cursor, first, n, n, aug, mean, std, p_hi, p_lo, lab)
    pos = torch.arange(first + 2 * n, first + 3 * n, device=DEV)
    idx = perm[pos]
    assert torch.equal(lab, labels[idx])
    imgs = []
    for j in range(n):
        img = data[idx[j]].permute(2, 0, 1).float()                     # PIL image -> CHW, still 0..255
        padded = F.pad(img, (4, 4, 4, 4))                               # RandomCrop(32, padding=4): black border
        dx, dy, fl = (int(v) for v in aug[pos[j], :3])
        crop = padded[:, dy:dy + 32, dx:dx + 32]
        if fl:
            crop = torch.flip(crop, dims=[2])                           # RandomHorizontalFlip
        t = crop / 255.0                                                # ToTensor
        t = (t - torch.tensor(mean, device=DEV)[:, None, None]) / torch.tensor(std, device=DEV)[:, None, None]
        imgs.append(t)
    xb = torch.stack(imgs)
    patches = F.unfold(xb, 3, padding=1).transpose(1, 2).reshape(n * 1024, 27)
    got = p_hi.float() + p_lo.float()
    assert float((got[:, :27] - patches).abs().max()) < 3e-5 * float(patches.abs().max())
    assert float(got[:, 27:].abs().max()) == 0.0
    # no augmentation (aug = None) = plain normalisation
    ops.stem_im2col_u8aug(data, labels, None, None, 0, 0, n, None, mean, std, p_hi, p_lo, lab)
    x0 = (data[:n].permute(0, 3, 1, 2).float() / 255.0 - torch.tensor(mean, device=DEV)[None, :, None, None]) / \
        torch.tensor(std, device=DEV)[None, :, None, None]
    ref0 = F.unfold(x0, 3, padding=1).transpose(1, 2).reshape(n * 1024, 27)
    assert float(((p_hi.float() + p_lo.float())[:, :27] - ref0).abs().max()) < 3e-5 * float(ref0.abs().max())


def test_unsupported_shapes_fail_loudly():
    """no fallback: a microbatch that the 128-pixel boxes cannot tile per group is refused"""
    with pytest.raises(RuntimeError):
        ops.tiles_per_group(12, 4, (4, 4, 8))
    a = L.ConvGemmArgs()
    rc = L.load().fb_conv_gemm(C.byref(a), None)
    assert rc != 0 and "fb_conv_gemm" in L.last_error()
