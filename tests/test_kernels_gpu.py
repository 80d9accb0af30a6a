"""Per-kernel parity tests on the B200: every exported kernel of the C ABI against plain torch ops (fp64) on the same
seeded inputs.  Tolerances are stated per test: the split (hi+lo bf16) tensor-core path carries ~16 mantissa bits per
operand, so products are exact to ~2^-16 relative and results are checked at 3e-5 of the output scale; the plain-bf16
path is checked against torch run on bf16-rounded operands (products then exact in fp32)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from fullbatchtraining_b200 import ops  # noqa: E402

DEV = torch.device("cuda")


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


def make_conv(n, h, w, cin, cout, k, stride, use_split, seed=0, dx=True, dx_accumulate=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(n, cin, h, w, device=DEV, generator=g)
    wt = torch.randn(cout, cin, k, k, device=DEV, generator=g) * (2.0 / (cout * k * k)) ** 0.5
    ho, wo = h // stride, w // stride
    gy = torch.randn(n, cout, ho, wo, device=DEV, generator=g) * 1e-3
    x_hi, x_lo = split(nhwc(x))
    if not use_split:
        x_lo = None
    taps = k * k
    wf_hi = torch.zeros(cout, taps * cin, device=DEV, dtype=torch.bfloat16)
    wf_lo = torch.zeros_like(wf_hi) if use_split else None
    wd_hi = torch.zeros(cin, taps * cout, device=DEV, dtype=torch.bfloat16)
    wd_lo = torch.zeros_like(wd_hi) if use_split else None
    ops.weight_prep(wt, cout, cin, taps, wf_hi, wf_lo, wd_hi, wd_lo)
    y = torch.full((n, ho, wo, cout), float("nan"), device=DEV)
    dy = nhwc(gy).to(torch.bfloat16)
    dxb = torch.full((n, h, w, cin), 0.5 if dx_accumulate else float("nan"), device=DEV) if dx else None
    partial = torch.empty(ops.Conv2dPlan.partial_elems(n, h, w, cin, cout, k, stride), device=DEV)
    plan = ops.Conv2dPlan(n, h, w, cin, cout, k, stride, x_hi, x_lo, y, dy, dxb, wf_hi, wf_lo, wd_hi, wd_lo, partial,
                          dx_accumulate=dx_accumulate, split=use_split)
    return dict(x=x, w=wt, gy=gy, plan=plan, y=y, dy=dy, dx=dxb, x_hi=x_hi, x_lo=x_lo, wf_hi=wf_hi, wf_lo=wf_lo)


CONV_CASES = [
    # n, h, w, cin, cout, k, stride
    (4, 32, 32, 64, 64, 3, 1),
    (2, 32, 32, 64, 128, 3, 2),
    (4, 16, 16, 128, 128, 3, 1),
    (8, 8, 8, 128, 256, 1, 1),
    (8, 16, 16, 128, 256, 3, 2),
    (16, 4, 4, 256, 512, 3, 1),
    (4, 4, 4, 512, 512, 3, 1),      # tile_n = 8 > n: out-of-bounds images are zero-filled and masked
    (16, 8, 8, 256, 512, 3, 2),
    (32, 8, 8, 64, 256, 1, 1),      # bottleneck 1x1 expansion, N tile 256
    (2, 32, 32, 128, 64, 3, 1),     # haloed-box kernel, 2 channel blocks
    (3, 16, 16, 64, 256, 3, 1),     # haloed-box kernel, 2 N tiles, odd image count
    (8, 8, 8, 256, 256, 3, 1),      # haloed-box kernel on 8x8 maps: two interleaved images per 128-pixel half
    (6, 8, 8, 128, 64, 3, 1),       # 8x8, image count only divisible by 2: single-half tiles
    (32, 4, 4, 128, 128, 3, 1),     # 4x4 maps: eight interleaved images per half
]


@pytest.fixture(params=["halo", "generic"], autouse=True)
def conv_path(request, monkeypatch):
    """3x3/stride-1 convs run through the generic per-tap kernel by default; the haloed-box kernel (FB_HALO=1) must stay
    correct for the same shapes.  The same switch selects the haloed (default) / per-tap wgrad variant."""
    if request.param == "halo":
        if "conv" not in request.node.name:
            pytest.skip("only conv tests depend on the conv path")
        monkeypatch.setenv("FB_HALO", "1")
        monkeypatch.setenv("FB_WGRAD_HALO", "1")
    else:
        monkeypatch.delenv("FB_HALO", raising=False)
        monkeypatch.setenv("FB_WGRAD_HALO", "0")
    return request.param


def operands(c, use_split):
    """The values the kernel really multiplies: split -> (hi+lo) of each operand (lo*lo term dropped, below tolerance)."""
    if use_split:
        return c["x"].double(), c["w"].double()
    return c["x"].to(torch.bfloat16).double(), c["w"].to(torch.bfloat16).double()


@pytest.mark.parametrize("use_split", [True, False], ids=["split", "bf16"])
@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_forward(case, use_split):
    n, h, w, cin, cout, k, stride = case
    c = make_conv(n, h, w, cin, cout, k, stride, use_split)
    c["plan"].forward()
    torch.cuda.synchronize()
    xr, wr = operands(c, use_split)
    ref = nhwc(F.conv2d(xr, wr, None, stride, (k - 1) // 2))
    assert torch.isfinite(c["y"]).all()
    assert rel_err(c["y"], ref) < 3e-5


@pytest.mark.parametrize("use_split", [True, False], ids=["split", "bf16"])
@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_dgrad(case, use_split):
    n, h, w, cin, cout, k, stride = case
    c = make_conv(n, h, w, cin, cout, k, stride, use_split)
    c["plan"].dgrad()
    torch.cuda.synchronize()
    _, wr = operands(c, use_split)
    gy = nchw(c["dy"].double())  # the kernel consumes the bf16-rounded output gradient
    ref = nhwc(torch.nn.grad.conv2d_input((n, cin, h, w), wr, gy, stride, (k - 1) // 2))
    assert torch.isfinite(c["dx"]).all()
    assert rel_err(c["dx"], ref) < 3e-5


@pytest.mark.parametrize("geom", [(8, 8, 8, 2, 2, 64), (8, 8, 8, 2, 1, 64), (8, 8, 8, 2, 2, 128), (8, 8, 8, 2, 1, 128),
                                  (16, 4, 4, 8, 1, 64), (16, 4, 4, 8, 2, 64), (16, 4, 4, 8, 2, 128),
                                  (4, 16, 16, 1, 1, 128), (4, 16, 16, 1, 2, 64), (2, 32, 32, 1, 1, 64)],
                         ids=lambda g: "x".join(map(str, g)))
@pytest.mark.parametrize("planes", [2, 1], ids=["split", "bf16"])
def test_conv3x3_tile_geometries(geom, planes, conv_path):
    """Every (imgs, halves, n_tile) geometry of fb_conv3x3, forced explicitly (the cost model picks one per layer)."""
    if conv_path == "generic":
        pytest.skip("geometry test drives fb_conv3x3 directly")
    n, h, w, imgs, halves, n_tile = geom
    if planes == 2 and (imgs, halves, n_tile) in ((2, 2, 128), (8, 2, 64), (8, 2, 128)):
        pytest.skip("two split halo boxes per stage plus the weight ring exceed 227 KB of shared memory")
    cin, cout = 128, 128
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(n, cin, h, w, device=DEV, generator=g)
    wt = torch.randn(cout, cin, 3, 3, device=DEV, generator=g) * (2.0 / (cout * 9)) ** 0.5
    x_hi, x_lo = split(nhwc(x))
    wf_hi = torch.zeros(cout, 9 * cin, device=DEV, dtype=torch.bfloat16)
    wf_lo = torch.zeros_like(wf_hi)
    ops.weight_prep(wt, cout, cin, 9, wf_hi, wf_lo)
    y = torch.full((n, h, w, cout), float("nan"), device=DEV)
    fk0 = [[(dhi * 3 + dwi) * cin for dhi in range(3)] for dwi in range(3)]
    conv = ops.Conv3x3([x_hi, x_lo][:planes], [wf_hi, wf_lo][:planes], n, h, w, cin, cout, fk0, y, False,
                       (imgs, halves, n_tile))
    conv()
    torch.cuda.synchronize()
    if planes == 2:
        xr, wr = x.double(), wt.double()
    else:
        xr, wr = x.to(torch.bfloat16).double(), wt.to(torch.bfloat16).double()
    ref = nhwc(F.conv2d(xr, wr, None, 1, 1))
    assert torch.isfinite(y).all()
    assert rel_err(y, ref) < 3e-5


def test_conv_dgrad_accumulate():
    n, h, w, cin, cout, k, stride = 2, 32, 32, 64, 128, 3, 2
    c = make_conv(n, h, w, cin, cout, k, stride, True, dx_accumulate=True)
    c["plan"].dgrad()
    torch.cuda.synchronize()
    gy = nchw(c["dy"].double())
    ref = nhwc(torch.nn.grad.conv2d_input((n, cin, h, w), c["w"].double(), gy, stride, 1)) + 0.5
    assert rel_err(c["dx"], ref) < 3e-5


@pytest.mark.parametrize("use_split", [True, False], ids=["split", "bf16"])
@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_wgrad(case, use_split):
    n, h, w, cin, cout, k, stride = case
    c = make_conv(n, h, w, cin, cout, k, stride, use_split, dx=False)
    g = torch.full((cout, cin, k, k), float("nan"), device=DEV)
    c["plan"].wgrad(g)
    torch.cuda.synchronize()
    xr, _ = operands(c, use_split)
    gy = nchw(c["dy"].double())
    ref = torch.nn.grad.conv2d_weight(xr, (cout, cin, k, k), gy, stride, (k - 1) // 2)
    assert torch.isfinite(g).all()
    assert rel_err(g, ref) < 3e-5


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


def test_stem_im2col_and_stem_conv():
    g = torch.Generator(device="cuda").manual_seed(5)
    n = 8
    data = torch.randn(32, 3, 32, 32, device=DEV, generator=g)
    labels = torch.randint(0, 10, (32,), device=DEV, generator=g)
    perm = torch.randperm(32, device=DEV, generator=g)
    cursor = torch.tensor([1], device=DEV, dtype=torch.int32)
    p_hi = torch.empty(n * 1024, 64, device=DEV, dtype=torch.bfloat16)
    p_lo = torch.empty_like(p_hi)
    lab = torch.empty(n, device=DEV, dtype=torch.int64)
    ops.stem_im2col(data, labels, perm, cursor, 8, n, p_hi, p_lo, lab)  # samples perm[8 + 1*8 : 8 + 2*8]
    idx = perm[16:24]
    x = data[idx]
    assert torch.equal(lab, labels[idx])
    patches = F.unfold(x, 3, padding=1).transpose(1, 2).reshape(n * 1024, 27)  # column = ci*9 + kh*3 + kw
    got = p_hi.float() + p_lo.float()
    assert float((got[:, :27] - patches).abs().max()) < 2e-5 * float(patches.abs().max())
    assert float(got[:, 27:].abs().max()) == 0.0
    # stem conv = 1x1 GEMM over the patches with the OIHW weight as B (27 of 64 columns used)
    w = torch.randn(64, 3, 3, 3, device=DEV, generator=g) * 0.1
    wf_hi = torch.zeros(64, 64, device=DEV, dtype=torch.bfloat16)
    wf_lo = torch.zeros_like(wf_hi)
    ops.weight_prep(w, 64, 3, 9, wf_hi, wf_lo)
    y = torch.empty(n, 32, 32, 64, device=DEV)
    gy = torch.randn(n, 64, 32, 32, device=DEV, generator=g) * 1e-3
    dy = nhwc(gy).to(torch.bfloat16)
    partial = torch.empty(ops.Conv2dPlan.partial_elems(n, 32, 32, 64, 64, 1, 1), device=DEV)
    plan = ops.Conv2dPlan(n, 32, 32, 64, 64, 1, 1, p_hi.view(n, 32, 32, 64), p_lo.view(n, 32, 32, 64), y, dy, None, wf_hi,
                          wf_lo, None, None, partial)
    plan.forward()
    ref = nhwc(F.conv2d(x.double(), w.double(), None, 1, 1))
    assert rel_err(y, ref) < 3e-5
    gw = torch.full((64, 3, 3, 3), float("nan"), device=DEV)
    plan.wgrad(gw, cin_real=3, mode=1)
    refw = torch.nn.grad.conv2d_weight(x.double(), (64, 3, 3, 3), nchw(dy.double()), 1, 1)
    assert rel_err(gw, refw) < 3e-5


@pytest.mark.parametrize("P,Cc", [(4096, 64), (131072, 64), (2048, 512), (512, 2048), (1000, 128)])
def test_bn_stats(P, Cc):
    g = torch.Generator(device="cuda").manual_seed(P + Cc)
    y = torch.randn(P, Cc, device=DEV, generator=g) * 2 + 0.5
    ws = torch.zeros(2 * Cc * 1024, device=DEV)  # zero-initialised: holds the block ticket
    mean, rstd = torch.empty(Cc, device=DEV), torch.empty(Cc, device=DEV)
    rm, rv = torch.zeros(Cc, device=DEV), torch.ones(Cc, device=DEV)
    ops.bn_stats(y, P, Cc, ws, mean, rstd, rm, rv)
    yd = y.double()
    m, v = yd.mean(0), yd.var(0, unbiased=False)
    assert rel_err(mean, m) < 1e-6
    assert rel_err(rstd, 1 / (v + 1e-5).sqrt()) < 1e-6
    assert rel_err(rm, 0.1 * m) < 1e-6
    assert rel_err(rv, 0.9 + 0.1 * yd.var(0, unbiased=True)) < 1e-6


@pytest.mark.parametrize("variant", ["plain", "residual", "dual", "norelu"])
def test_bn_apply_and_backward(variant):
    g = torch.Generator(device="cuda").manual_seed(11)
    n, h, w, Cc = 4, 8, 8, 128
    P = n * h * w
    y = (torch.randn(P, Cc, device=DEV, generator=g) * 1.5 + 0.3).requires_grad_(False)
    gamma = torch.rand(Cc, device=DEV, generator=g) + 0.5
    beta = torch.randn(Cc, device=DEV, generator=g) * 0.1
    y2 = torch.randn(P, Cc, device=DEV, generator=g)
    gamma2 = torch.rand(Cc, device=DEV, generator=g) + 0.5
    beta2 = torch.randn(Cc, device=DEV, generator=g) * 0.1
    res = torch.randn(P, Cc, device=DEV, generator=g)
    res_hi, res_lo = split(res)
    ws = torch.zeros(2 * Cc * 1024, device=DEV)  # zero-initialised: holds the block ticket
    mean, rstd = torch.empty(Cc, device=DEV), torch.empty(Cc, device=DEV)
    mean2, rstd2 = torch.empty(Cc, device=DEV), torch.empty(Cc, device=DEV)
    ops.bn_stats(y, P, Cc, ws, mean, rstd, None, None)
    ops.bn_stats(y2, P, Cc, ws, mean2, rstd2, None, None)
    out_hi = torch.empty(P, Cc, device=DEV, dtype=torch.bfloat16)
    out_lo = torch.empty_like(out_hi)
    relu = variant != "norelu"
    ops.bn_apply(y, mean, rstd, gamma, beta, P, Cc, out_hi, out_lo, relu=relu,
                 second=(y2, mean2, rstd2, gamma2, beta2) if variant == "dual" else None,
                 res=(res_hi, res_lo) if variant == "residual" else None)
    # fp64 reference through autograd
    yd = y.double().requires_grad_(True)
    y2d = y2.double().requires_grad_(True)
    resd = (res_hi.double() + res_lo.double()).requires_grad_(True)

    def bn(t, ga, be):
        t4 = t.view(n, h, w, Cc).permute(0, 3, 1, 2)
        o = F.batch_norm(t4, None, None, ga.double(), be.double(), True, 0.1, 1e-5)
        return o.permute(0, 2, 3, 1).reshape(P, Cc)

    ref = bn(yd, gamma, beta)
    if variant == "dual":
        ref = ref + bn(y2d, gamma2, beta2)
    if variant == "residual":
        ref = ref + resd
    if relu:
        ref = F.relu(ref)
    got = out_hi.double() + out_lo.double()
    assert rel_err(got, ref) < 2e-5
    # backward
    dA = torch.randn(P, Cc, device=DEV, generator=g)
    gy_ref, = torch.autograd.grad(ref, yd, dA.double(), retain_graph=True)
    dgamma, dbeta = torch.empty(Cc, device=DEV), torch.empty(Cc, device=DEV)
    dy = torch.empty(P, Cc, device=DEV, dtype=torch.bfloat16)
    dz = torch.empty(P, Cc, device=DEV)
    ops.bn_bwd(dA, out_hi if relu else None, y, mean, rstd, gamma, P, Cc, ws, dgamma, dbeta, dy, dz_out=dz)
    # dy is stored in bf16 (8 mantissa bits): 2^-8 relative to the output scale
    assert rel_err(dy, gy_ref) < 5e-3
    mask = (ref > 0).double() if relu else torch.ones_like(ref)
    assert rel_err(dz, dA.double() * mask) < 1e-6
    xhat = (yd - yd.mean(0)) / (yd.var(0, unbiased=False) + 1e-5).sqrt()
    assert rel_err(dgamma, (dA.double() * mask * xhat).sum(0)) < 2e-5
    assert rel_err(dbeta, (dA.double() * mask).sum(0)) < 2e-5
    if variant == "residual":
        gres, = torch.autograd.grad(ref, resd, dA.double())
        assert rel_err(dz, gres) < 1e-6


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
def test_head(smoothing):
    g = torch.Generator(device="cuda").manual_seed(9)
    n, hw, c, classes = 32, 16, 512, 10
    a = torch.relu(torch.randn(n, hw, c, device=DEV, generator=g))
    hi, lo = split(a)
    fc_w = torch.randn(classes, c, device=DEV, generator=g) * 0.05
    fc_b = torch.randn(classes, device=DEV, generator=g) * 0.1
    labels = torch.randint(0, classes, (n,), device=DEV, generator=g)
    ws = torch.empty(n * (c + 32), device=DEV)
    scal = torch.zeros(16, device=DEV)
    d_w, d_b = torch.empty_like(fc_w), torch.empty_like(fc_b)
    dA = torch.empty(n, hw, c, device=DEV)
    ops.head_fwd_bwd(hi, lo, n, hw, c, fc_w, fc_b, labels, classes, smoothing, ws, scal, 2, 3, d_w, d_b, dA)
    ad = (hi.double() + lo.double()).requires_grad_(True)
    wd, bd = fc_w.double().requires_grad_(True), fc_b.double().requires_grad_(True)
    logits = F.linear(ad.mean(1), wd, bd)
    logp = F.log_softmax(logits, -1)
    wgt = torch.full_like(logits, smoothing / (classes - 1))
    wgt.scatter_(-1, labels.unsqueeze(-1), 1 - smoothing)
    loss = (-wgt * logp).sum(-1).mean()
    ga, gw, gb = torch.autograd.grad(loss, (ad, wd, bd))
    assert abs(float(scal[2]) - float(loss)) < 1e-5 * abs(float(loss))
    assert float(scal[3]) == float((logits.argmax(-1) == labels).sum())
    assert rel_err(dA, ga) < 1e-5
    assert rel_err(d_w, gw) < 1e-5
    assert rel_err(d_b, gb) < 1e-5


def test_flat_fd_kernels():
    g = torch.Generator(device="cuda").manual_seed(4)
    n = 1_000_003  # not a multiple of 4: exercises the tails
    theta = torch.randn(n + 1, device=DEV, generator=g)[:n]
    grad = torch.randn(n + 1, device=DEV, generator=g)[:n] * 1e-3
    g2 = grad + torch.randn(n, device=DEV, generator=g) * 1e-6
    avg = torch.randn(n, device=DEV, generator=g) * 1e-3
    ws = torch.empty(1024, device=DEV, dtype=torch.float64)
    scal = torch.zeros(16, device=DEV)
    cursor = torch.tensor([2], device=DEV, dtype=torch.int32)
    norms = torch.zeros(8, device=DEV)
    ops.flat_sqnorm(grad, n, ws, scal, 0, norms, cursor)
    n2 = float(grad.double().pow(2).sum())
    assert abs(float(scal[0]) - n2) < 1e-6 * n2
    theta_p = torch.empty_like(theta)
    bs, eps, cf = 0.5, 1e-2, 0.2
    ops.fd_perturb(theta, grad, n, bs, eps, scal, 0, 1, theta_p)
    eps_n = eps / (bs * n2 ** 0.5)
    assert abs(float(scal[1]) - eps_n) < 1e-6 * eps_n
    assert float(norms[2]) == float(scal[0])
    assert rel_err(theta_p, theta.double() + eps_n * bs * grad.double()) < 1e-6
    g_in, avg_in = grad.clone(), avg.clone()
    ops.fd_combine(g_in, g2, avg_in, n, scal, 1, cf, cursor, 4, True)
    en = float(scal[1])
    g_reg = grad.double() + cf * (g2.double() - grad.double()) / en
    # (g2 - g) is formed in fp32 from nearly equal numbers: one fp32 ulp of g, amplified by cf/eps_n
    tol = float(grad.abs().max()) * 2 ** -23 * cf / en * 2
    assert float((g_in.double() - g_reg).abs().max()) < tol
    ref_avg = avg.double() + (g_in.double() - avg.double()) / 7
    assert rel_err(avg_in, ref_avg) < 1e-6
    avg2 = avg.clone()
    ops.mean_accumulate(grad, avg2, n, None, 0)
    assert rel_err(avg2, grad) < 1e-6
    ops.cursor_add(cursor, 3)
    assert int(cursor) == 5
    x = grad.clone()
    ops.flat_scale(x, n, 0.25)
    assert torch.equal(x, grad * 0.25)


def test_bn_bwd_second_addend():
    """dA2: the shortcut-branch gradient is added on the fly (no read-modify-write in a GEMM epilogue)."""
    g = torch.Generator(device="cuda").manual_seed(21)
    P, Cc = 2048, 64
    y = torch.randn(P, Cc, device=DEV, generator=g)
    gamma = torch.rand(Cc, device=DEV, generator=g) + 0.5
    a1 = torch.randn(P, Cc, device=DEV, generator=g)
    a2 = torch.randn(P, Cc, device=DEV, generator=g)
    mask = torch.randn(P, Cc, device=DEV, generator=g).to(torch.bfloat16)
    ws = torch.zeros(2 * Cc * 1024, device=DEV)
    mean, rstd = torch.empty(Cc, device=DEV), torch.empty(Cc, device=DEV)
    ops.bn_stats(y, P, Cc, ws, mean, rstd, None, None)
    outs = []
    for dA, dA2 in ((a1, a2), (a1 + a2, None)):
        dgamma, dbeta = torch.empty(Cc, device=DEV), torch.empty(Cc, device=DEV)
        dy = torch.empty(P, Cc, device=DEV, dtype=torch.bfloat16)
        dz = torch.empty(P, Cc, device=DEV)
        ops.bn_bwd(dA, mask, y, mean, rstd, gamma, P, Cc, ws, dgamma, dbeta, dy, dz_out=dz, dA2=dA2)
        outs.append((dgamma, dbeta, dy.float(), dz))
    for u, v in zip(*outs):
        assert torch.equal(u, v)


@pytest.mark.parametrize("P,Cc,variant", [(131072, 64, "residual"), (32768, 128, "dual"), (2048, 512, "plain"),
                                          (512, 2048, "plain"), (1000, 128, "residual"), (8192, 256, "dual")])
@pytest.mark.parametrize("sliced", [False, True], ids=["barrier", "cluster"])
def test_bn_fused_kernels_match_unfused(P, Cc, variant, sliced, monkeypatch):
    """fb_bn_fwd_fused / fb_bn_bwd_fused (one persistent launch, grid barriers; or the opt-in small-map variant with a
    thread-block cluster + distributed shared memory, FB_BN_SLICED_MAX) against the three-launch kernels."""
    if sliced:
        if P * Cc > 4500000:
            pytest.skip("small-map variant")
        monkeypatch.setenv("FB_BN_SLICED_MAX", "4500000")
    else:
        monkeypatch.delenv("FB_BN_SLICED_MAX", raising=False)
    g = torch.Generator(device="cuda").manual_seed(P + Cc)
    y = torch.randn(P, Cc, device=DEV, generator=g) * 1.5 + 0.3
    y2 = torch.randn(P, Cc, device=DEV, generator=g)
    gamma = torch.rand(Cc, device=DEV, generator=g) + 0.5
    beta = torch.randn(Cc, device=DEV, generator=g) * 0.1
    gamma2 = torch.rand(Cc, device=DEV, generator=g) + 0.5
    beta2 = torch.randn(Cc, device=DEV, generator=g) * 0.1
    res_hi, res_lo = split(torch.randn(P, Cc, device=DEV, generator=g))
    ws = torch.zeros(2 * Cc * 1024, device=DEV)
    ws2 = torch.zeros(2 * Cc * 1024, device=DEV)
    # reference: unfused kernels
    mean, rstd, mean2, rstd2 = (torch.empty(Cc, device=DEV) for _ in range(4))
    rm, rv = torch.zeros(Cc, device=DEV), torch.ones(Cc, device=DEV)
    ops.bn_stats(y, P, Cc, ws, mean, rstd, rm, rv)
    ops.bn_stats(y2, P, Cc, ws, mean2, rstd2, None, None)
    o_hi, o_lo = torch.empty(P, Cc, device=DEV, dtype=torch.bfloat16), torch.empty(P, Cc, device=DEV, dtype=torch.bfloat16)
    ops.bn_apply(y, mean, rstd, gamma, beta, P, Cc, o_hi, o_lo, relu=True,
                 second=(y2, mean2, rstd2, gamma2, beta2) if variant == "dual" else None,
                 res=(res_hi, res_lo) if variant == "residual" else None)
    # fused
    fmean, frstd, fmean2, frstd2 = (torch.empty(Cc, device=DEV) for _ in range(4))
    frm, frv = torch.zeros(Cc, device=DEV), torch.ones(Cc, device=DEV)
    frm2, frv2 = torch.zeros(Cc, device=DEV), torch.ones(Cc, device=DEV)
    f_hi, f_lo = torch.empty_like(o_hi), torch.empty_like(o_lo)
    for _ in range(2):  # twice: the barrier counters must reset themselves
        frm.zero_(); frv.fill_(1)
        ops.bn_fwd_fused(y, fmean, frstd, gamma, beta, P, Cc, f_hi, f_lo, ws2, running=(frm, frv), relu=True,
                         second=(y2, fmean2, frstd2, gamma2, beta2, frm2, frv2) if variant == "dual" else None,
                         res=(res_hi, res_lo) if variant == "residual" else None)
    assert rel_err(fmean, mean) < 1e-6 and rel_err(frstd, rstd) < 1e-6
    assert rel_err(frm, rm) < 1e-6 and rel_err(frv, rv) < 1e-6
    got, ref = f_hi.double() + f_lo.double(), o_hi.double() + o_lo.double()
    assert rel_err(got, ref) < 2e-5
    # backward
    dA = torch.randn(P, Cc, device=DEV, generator=g)
    dA2 = torch.randn(P, Cc, device=DEV, generator=g)
    outs = []
    for fused in (False, True):
        dgamma, dbeta = torch.empty(Cc, device=DEV), torch.empty(Cc, device=DEV)
        dy = torch.empty(P, Cc, device=DEV, dtype=torch.bfloat16)
        dz = torch.empty(P, Cc, device=DEV)
        fn = ops.bn_bwd_fused if fused else ops.bn_bwd
        for _ in range(2 if fused else 1):
            fn(dA, o_hi, y, mean, rstd, gamma, P, Cc, ws2 if fused else ws, dgamma, dbeta, dy, dz_out=dz, dA2=dA2)
        outs.append((dgamma, dbeta, dy.float(), dz))
    assert rel_err(outs[1][0], outs[0][0]) < 1e-5 and rel_err(outs[1][1], outs[0][1]) < 1e-5
    assert rel_err(outs[1][2], outs[0][2]) < 1e-2   # bf16 outputs: last-bit differences from the reduction order
    assert torch.equal(outs[1][3], outs[0][3])


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
    ops.stem_im2col_u8aug(data, labels, perm, cursor, first, n, aug, mean, std, p_hi, p_lo, lab)
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
    ops.stem_im2col_u8aug(data, labels, None, None, 0, n, None, mean, std, p_hi, p_lo, lab)
    x0 = (data[:n].permute(0, 3, 1, 2).float() / 255.0 - torch.tensor(mean, device=DEV)[None, :, None, None]) / \
        torch.tensor(std, device=DEV)[None, :, None, None]
    ref0 = F.unfold(x0, 3, padding=1).transpose(1, 2).reshape(n * 1024, 27)
    assert float(((p_hi.float() + p_lo.float())[:, :27] - ref0).abs().max()) < 3e-5 * float(ref0.abs().max())


@pytest.mark.parametrize("case", [(4, 32, 32, 64, 64, 3, 1), (3, 16, 16, 64, 256, 3, 1), (2, 32, 32, 64, 128, 3, 2),
                                  (16, 8, 8, 128, 256, 3, 1), (4, 4, 4, 512, 512, 3, 1), (32, 8, 8, 64, 256, 1, 1),
                                  (16, 4, 4, 256, 512, 3, 1)],
                         ids=lambda c: "x".join(map(str, c)))
def test_conv_forward_fused_statistics(case):
    """BatchNorm statistics fused into the conv epilogue: per-CTA column sums / sums of squares of the output."""
    n, h, w, cin, cout, k, stride = case
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, cin, h, w, device=DEV, generator=g)
    wt = torch.randn(cout, cin, k, k, device=DEV, generator=g) * (2.0 / (cout * k * k)) ** 0.5
    ho, wo = h // stride, w // stride
    x_hi, x_lo = split(nhwc(x))
    taps = k * k
    wf_hi = torch.zeros(cout, taps * cin, device=DEV, dtype=torch.bfloat16)
    wf_lo = torch.zeros_like(wf_hi)
    ops.weight_prep(wt, cout, cin, taps, wf_hi, wf_lo)
    y = torch.empty(n, ho, wo, cout, device=DEV)
    dy = torch.zeros(n, ho, wo, cout, device=DEV, dtype=torch.bfloat16)
    partial = torch.empty(ops.Conv2dPlan.partial_elems(n, h, w, cin, cout, k, stride), device=DEV)
    plan = ops.Conv2dPlan(n, h, w, cin, cout, k, stride, x_hi, x_lo, y, dy, None, wf_hi, wf_lo, None, None, partial,
                          fuse_stats=True)
    plan.forward()
    buf, rows = plan.stats
    assert buf.shape[0] == rows >= 1
    yd = y.double().reshape(-1, cout)
    s1, s2 = buf[:, 0].double().sum(0), buf[:, 1].double().sum(0)
    assert rel_err(s1, yd.sum(0)) < 1e-5
    assert rel_err(s2, (yd * yd).sum(0)) < 1e-5
    # and the fused BatchNorm consumes them: same result as with its own statistics pass
    P = n * ho * wo
    gamma, beta = torch.rand(cout, device=DEV, generator=g) + 0.5, torch.randn(cout, device=DEV, generator=g) * 0.1
    ws = torch.zeros(2 * cout * 1024, device=DEV)
    outs = []
    for stats in (None, plan.stats, "sliced"):
        if stats == "sliced":  # the barrier-free forward variant consumes the same epilogue statistics
            os.environ["FB_BN_SLICED_MAX"] = "4500000"
            stats = plan.stats
        mean, rstd = torch.empty(cout, device=DEV), torch.empty(cout, device=DEV)
        rm, rv = torch.zeros(cout, device=DEV), torch.ones(cout, device=DEV)
        hi = torch.empty(P, cout, device=DEV, dtype=torch.bfloat16)
        lo = torch.empty_like(hi)
        for _ in range(2):
            rm.zero_(); rv.fill_(1)
            ops.bn_fwd_fused(y, mean, rstd, gamma, beta, P, cout, hi, lo, ws, running=(rm, rv), stats=stats)
        outs.append((mean, rstd, rm, rv, hi.double() + lo.double()))
    os.environ.pop("FB_BN_SLICED_MAX", None)
    for other in outs[1:]:
        for u, v in zip(outs[0], other):
            assert rel_err(u, v) < 2e-5


@pytest.mark.parametrize("case", [(4, 32, 32, 64, 64, 3, 1), (8, 16, 16, 128, 128, 3, 1), (16, 8, 8, 256, 128, 1, 1),
                                  (16, 4, 4, 512, 512, 3, 1)], ids=lambda c: "x".join(map(str, c)))
def test_dgrad_epilogue_bn_backward_statistics(case, conv_path):
    """BatchNorm-backward statistics fused into the dgrad epilogue (fb_conv_gemm_args.bwd_y): sums of dA*m and
    dA*m*xhat of the stored gradient, and fb_bn_bwd_fused consuming them gives what its own first pass gives."""
    if conv_path == "halo":
        pytest.skip("the haloed kernel has no backward-statistics epilogue")
    n, h, w, cin, cout, k, stride = case
    g = torch.Generator(device="cuda").manual_seed(5)
    wt = torch.randn(cout, cin, k, k, device=DEV, generator=g) * (2.0 / (cout * k * k)) ** 0.5
    gy = torch.randn(n, cout, h, w, device=DEV, generator=g) * 1e-2
    x_hi, x_lo = split(torch.randn(n, h, w, cin, device=DEV, generator=g))
    taps = k * k
    wf_hi = torch.zeros(cout, taps * cin, device=DEV, dtype=torch.bfloat16)
    wf_lo, wd_hi = torch.zeros_like(wf_hi), torch.zeros(cin, taps * cout, device=DEV, dtype=torch.bfloat16)
    wd_lo = torch.zeros_like(wd_hi)
    ops.weight_prep(wt, cout, cin, taps, wf_hi, wf_lo, wd_hi, wd_lo)
    y = torch.empty(n, h, w, cout, device=DEV)
    dy = nhwc(gy).to(torch.bfloat16)
    dx = torch.full((n, h, w, cin), float("nan"), device=DEV)
    # the BatchNorm(+ReLU) that produced the conv input: its pre-activation, output plane, mean, rstd
    P = n * h * w
    by = torch.randn(P, cin, device=DEV, generator=g) * 1.3 + 0.2
    bmean, bvar = by.mean(0), by.var(0, unbiased=False)
    brstd = torch.rsqrt(bvar + 1e-5)
    gamma = torch.rand(cin, device=DEV, generator=g) + 0.5
    act = torch.relu((by - bmean) * brstd * gamma + 0.1 * torch.randn(cin, device=DEV, generator=g))
    mask_hi = act.to(torch.bfloat16)
    partial = torch.empty(ops.Conv2dPlan.partial_elems(n, h, w, cin, cout, k, stride), device=DEV)
    plan = ops.Conv2dPlan(n, h, w, cin, cout, k, stride, x_hi, x_lo, y, dy, dx, wf_hi, wf_lo, wd_hi, wd_lo, partial,
                          dgrad_bn=(by, mask_hi, bmean, brstd))
    assert plan.dgrad_stats is not None
    plan.dgrad()
    buf, rows = plan.dgrad_stats
    d = dx.double().reshape(P, cin)
    m = (mask_hi.double() > 0).double()
    xhat = (by.double() - bmean.double()) * brstd.double()
    s1, s2 = buf[:, 0].double().sum(0), buf[:, 1].double().sum(0)
    assert rel_err(s1, (d * m).sum(0)) < 1e-5
    assert rel_err(s2, (d * m * xhat).sum(0)) < 1e-5
    ws = torch.zeros(2 * cin * 1024, device=DEV)
    outs = []
    for stats in (None, plan.dgrad_stats):
        dg, db = torch.empty(cin, device=DEV), torch.empty(cin, device=DEV)
        dyo = torch.empty(P, cin, device=DEV, dtype=torch.bfloat16)
        ops.bn_bwd_fused(dx, mask_hi, by, bmean, brstd, gamma, P, cin, ws, dg, db, dyo, stats=stats)
        outs.append((dg, db, dyo.double()))
    for u, v in zip(*outs):
        assert rel_err(u, v) < 2e-4
