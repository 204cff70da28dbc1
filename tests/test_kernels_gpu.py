"""GPU parity of the bandwidth-bound kernels (norm/act, SPFF tail, pooling, stem, transposed conv,
head, loss, Adam) against the oracle's / PyTorch's fp32 formulas on the same inputs.
Tolerances: tensors stored in bf16 -> rel-L2 <= 5e-3 (8-bit mantissa); fp32 reductions -> 1e-4."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def pm(x, ld=None):  # NCDHW fp32 -> position-major bf16
    n, c, d, h, w = x.shape
    ld = ld or c
    buf = torch.zeros(n, d, h, w, ld, dtype=torch.bfloat16, device=x.device)
    buf[..., :c] = x.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
    return buf[..., :c] if ld != c else buf


def ncdhw(buf):
    return buf.permute(0, 4, 1, 2, 3).float()


def _coef(xb, c, gamma, beta, eps=1e-5):
    from spff_b200 import ops
    n, d, h, w, _ = xb.shape
    stats = torch.zeros(n, c, 2, dtype=torch.float64, device="cuda")
    ops.in_stats(xb, c, stats)
    coef = torch.empty(n, c, 4, device="cuda")
    ops.in_coeffs(stats, gamma, beta, eps, n, c, d * h * w, coef)
    return coef


@pytest.mark.parametrize("n,c,d,h,w", [(2, 32, 5, 16, 16), (1, 64, 5, 8, 8), (2, 256, 5, 2, 2), (1, 128, 3, 6, 10)])
def test_instnorm_lrelu(n, c, d, h, w):
    from spff_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(n, c, d, h, w, device="cuda") * 2 + 0.5
    gamma = torch.rand(c, device="cuda") + 0.5
    beta = torch.randn(c, device="cuda") * 0.1
    xb = pm(x)
    coef = _coef(xb, c, gamma, beta)
    y = torch.empty_like(xb)
    ops.norm_act_apply(xb, coef, y, c, 0.01)
    xr = ncdhw(xb)
    ref = F.leaky_relu(F.instance_norm(xr, weight=gamma, bias=beta, eps=1e-5), 0.01)
    assert rel(ncdhw(y), ref) < 5e-3
    mean = xr.mean(dim=(2, 3, 4))
    assert torch.allclose(coef[..., 2], mean, atol=1e-4)
    S = torch.zeros(n, d, c, device="cuda")
    ops.norm_act_reduce(xb, coef, S, c, 0.01)
    assert rel(S, ref.sum(dim=(3, 4)).permute(0, 2, 1)) < 1e-3


def _tail_params(c, hid, d, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g).cuda()
    return dict(
        gamma=(1 + 0.1 * r(c)).requires_grad_(), beta=(0.1 * r(c)).requires_grad_(),
        w0=(0.3 * r(32, 16, 1)).requires_grad_(), b0=(0.1 * r(32)).requires_grad_(),
        w2=(0.3 * r(2 * c, 32, 1)).requires_grad_(), b2=(0.1 * r(2 * c)).requires_grad_(),
        mask=(1 + 0.2 * r(1, 1, d // 2 + 1, 1, 1)).requires_grad_(), scale=(1 + 0.1 * r(1)).requires_grad_(),
        sw1=(0.3 * r(hid, c, 1, 1, 1)).requires_grad_(), sb1=(0.1 * r(hid)).requires_grad_(),
        sw2=(0.3 * r(c, hid, 1, 1, 1)).requires_grad_(), sb2=(0.1 * r(c)).requires_grad_(),
    )


def _ref_tail(x, p, flags):
    """reference order: IN -> LReLU -> EFiLM -> FourierGate -> SpectralSE -> ChannelSE (oracle formulas)."""
    from oracle import spff_oracle as O
    q = {"e.mlp.0.weight": p["w0"], "e.mlp.0.bias": p["b0"], "e.mlp.2.weight": p["w2"], "e.mlp.2.bias": p["b2"],
         "f.freq_mask": p["mask"], "f.mag_scale": p["scale"],
         "s.fc.0.weight": p["sw1"], "s.fc.0.bias": p["sb1"], "s.fc.2.weight": p["sw2"], "s.fc.2.bias": p["sb2"]}
    q = {k: v.cpu() for k, v in q.items()}
    a = F.leaky_relu(F.instance_norm(x, weight=p["gamma"].cpu(), bias=p["beta"].cpu(), eps=1e-5), 0.01)
    if flags & 1:
        a = O._efilm(q, "e", a)
    if flags & 2:
        a = O._fgate(q, "f", a)
    if flags & 4:
        a = O._spectral_se(a)
    if flags & 8:
        a = O._channel_se(q, "s", a)
    return a


@pytest.mark.parametrize("flags", [0, 1, 2, 3, 15, 12])
@pytest.mark.parametrize("n,c,d,h,w", [(2, 32, 5, 8, 8), (1, 64, 5, 4, 4)])
def test_spff_tail_fwd_bwd(flags, n, c, d, h, w):
    """out = lrelu(IN(x))*P+Q and its full backward (dx, dgamma, dbeta, gate parameter grads) against
    autograd through the reference-ordered ops (models.py:1473-1478, 684-685)."""
    from spff_b200 import ops, tables
    from spff_b200._lib import Shape
    hid = max(4, c // 16)
    p = _tail_params(c, hid, d, 3)
    torch.manual_seed(1)
    x = torch.randn(n, c, d, h, w, device="cuda") * 1.5 + 0.3
    xb = pm(x)
    shape = Shape(n, d, h, w)
    coef = _coef(xb, c, p["gamma"].detach(), p["beta"].detach())
    g1 = bt = kfg = None
    if flags & 1:
        g1, bt = tables.efilm_tables(p["w0"], p["b0"], p["w2"], p["b2"], c, d)
    if flags & 2:
        kfg = tables.fourier_kernel(p["mask"], p["scale"], d)
    se = None
    if flags & 8:
        se = (p["sw1"].detach().reshape(hid, c).contiguous(), p["sb1"].detach(), p["sw2"].detach().reshape(c, hid).contiguous(),
              p["sb2"].detach())
    dt = lambda t: t.detach().contiguous() if t is not None else None
    out = torch.empty_like(xb)
    P = Q = S = None
    if flags:
        S = torch.zeros(n, d, c, device="cuda")
        ops.norm_act_reduce(xb, coef, S, c, 0.01)
        P = torch.empty(n, d, c, device="cuda"); Q = torch.empty(n, d, c, device="cuda")
        ops.gate_micro_fwd(S, dt(g1), dt(bt), dt(kfg), se, flags, c, shape, P, Q)
    ops.norm_act_affine_apply(xb, coef, P, Q, out, None, c, 0.01)
    # reference (CPU fp32 autograd)
    xr = ncdhw(xb).cpu().requires_grad_()
    ref = _ref_tail(xr, p, flags)
    assert rel(ncdhw(out).cpu(), ref) < 5e-3
    # backward
    torch.manual_seed(2)
    dout = torch.randn(n, c, d, h, w, device="cuda")
    dob = pm(dout)
    ref.backward(ncdhw(dob).cpu())
    R = torch.zeros(n, d, c, 6, device="cuda")
    ops.norm_act_bwd_reduce(dob, xb, coef, R, c, 0.01)
    bcoef = torch.empty(n, c, 4, device="cuda")
    dSa = torch.empty(n, d, c, device="cuda") if flags else None
    Pout = torch.empty(n, d, c, device="cuda") if flags else None
    z = lambda *s: torch.zeros(*s, device="cuda")
    dgamma, dbeta = z(c), z(c)
    dg1 = z(c, d) if flags & 1 else None
    dbt = z(c, d) if flags & 1 else None
    dk = z(d) if flags & 2 else None
    dse = (z(hid, c), z(hid), z(c, hid), z(c)) if flags & 8 else None
    ops.gate_micro_bwd(R, S, coef, p["gamma"].detach(), dt(g1), dt(bt), dt(kfg), se, flags, c, shape, bcoef, dSa, Pout,
                       dgamma, dbeta, dg1, dbt, dk, dse)
    dx = torch.empty_like(xb)
    ops.norm_act_bwd_apply(dob, xb, coef, bcoef, Pout, dSa, dx, c, 0.01)
    torch.cuda.synchronize()
    assert rel(ncdhw(dx).cpu(), xr.grad) < 1e-2
    assert rel(dgamma.cpu(), p["gamma"].grad.cpu()) < 2e-3
    assert rel(dbeta.cpu(), p["beta"].grad.cpu()) < 2e-3
    if flags & 8:
        assert rel(dse[0].cpu(), p["sw1"].grad.reshape(hid, c).cpu()) < 5e-3
        assert rel(dse[3].cpu(), p["sb2"].grad.cpu()) < 5e-3


@pytest.mark.parametrize("n,c,d,h,w", [(2, 32, 5, 16, 16), (3, 64, 5, 24, 8), (2, 256, 5, 4, 4), (1, 128, 16, 8, 8), (1, 512, 1, 4, 4)])
def test_bwd_reduce_lean_path_matches_direct_sums(n, c, d, h, w):
    """Fixed-order path with the lean first stage (xhat sums derived from {dout*a, dout*m, m, S}) == the direct sums
    of the original kernel, for the full (gated) and the plain statistics, ragged chunk tails included."""
    from spff_b200 import ops
    torch.manual_seed(7)
    x = torch.randn(n, c, d, h, w, device="cuda") * 1.3 + 0.4
    gamma = torch.rand(c, device="cuda") + 0.5
    gamma[1] = -0.7
    beta = torch.randn(c, device="cuda") * 0.3
    xb = pm(x)
    dob = pm(torch.randn(n, c, d, h, w, device="cuda"))
    coef = _coef(xb, c, gamma, beta)
    S = torch.zeros(n, d, c, device="cuda")
    ops.norm_act_reduce(xb, coef, S, c, 0.01)
    direct = torch.zeros(n, d, c, 6, device="cuda")
    ops.norm_act_bwd_reduce(dob, xb, coef, direct, c, 0.01)
    scale = direct.abs().amax(dim=(0, 1, 2), keepdim=True) + 1e-6
    lean = torch.full((n, d, c, 6), float("nan"), device="cuda")
    ops.norm_act_bwd_reduce(dob, xb, coef, lean, c, 0.01, fixed_order=True, S=S)
    assert float(((lean - direct) / scale).abs().max()) < 2e-4
    lean2 = torch.full((n, d, c, 6), float("nan"), device="cuda")
    ops.norm_act_bwd_reduce(dob, xb, coef, lean2, c, 0.01, fixed_order=True, S=S)
    assert torch.equal(lean, lean2)                      # fixed order: bit-reproducible
    plain = torch.zeros(n, d, c, 6, device="cuda")
    ops.norm_act_bwd_reduce(dob, xb, coef, plain, c, 0.01, plain=True, fixed_order=True)
    assert float(((plain[..., [2, 4]] - direct[..., [2, 4]]) / scale[..., [2, 4]]).abs().max()) < 2e-4
    # ReLU (slope 0), as the 3DUNet control runs it
    direct0 = torch.zeros(n, d, c, 6, device="cuda")
    ops.norm_act_bwd_reduce(dob, xb, coef, direct0, c, 0.0, plain=True)
    plain0 = torch.zeros(n, d, c, 6, device="cuda")
    ops.norm_act_bwd_reduce(dob, xb, coef, plain0, c, 0.0, plain=True, fixed_order=True)
    assert float(((plain0[..., [2, 4]] - direct0[..., [2, 4]]) / scale[..., [2, 4]]).abs().max()) < 2e-4


def test_tail_table_grads():
    """Gradients reaching the EFiLM MLP and the FourierGate mask through the CUDA tables path."""
    from spff_b200 import ops, tables
    from spff_b200._lib import Shape
    n, c, d, h, w = 2, 32, 5, 8, 8
    flags = 3
    pa = _tail_params(c, 4, d, 5)
    pb = {k: v.detach().clone().requires_grad_() for k, v in pa.items()}
    torch.manual_seed(4)
    x = torch.randn(n, c, d, h, w, device="cuda") + 0.2
    xb = pm(x)
    dob = pm(torch.randn(n, c, d, h, w, device="cuda"))
    ref = _ref_tail(ncdhw(xb).cpu(), pb, flags)
    ref.backward(ncdhw(dob).cpu())
    shape = Shape(n, d, h, w)
    coef = _coef(xb, c, pa["gamma"].detach(), pa["beta"].detach())
    g1, bt = tables.efilm_tables(pa["w0"], pa["b0"], pa["w2"], pa["b2"], c, d)
    kfg = tables.fourier_kernel(pa["mask"], pa["scale"], d)
    S = torch.zeros(n, d, c, device="cuda")
    ops.norm_act_reduce(xb, coef, S, c, 0.01)
    R = torch.zeros(n, d, c, 6, device="cuda")
    ops.norm_act_bwd_reduce(dob, xb, coef, R, c, 0.01)
    z = lambda *s: torch.zeros(*s, device="cuda")
    bcoef, dSa, Pout = z(n, c, 4), z(n, d, c), z(n, d, c)
    dg1, dbt, dk = z(c, d), z(c, d), z(d)
    ops.gate_micro_bwd(R, S, coef, pa["gamma"].detach(), g1.detach(), bt.detach(), kfg.detach(), None, flags, c, shape,
                       bcoef, dSa, Pout, z(c), z(c), dg1, dbt, dk, None)
    torch.autograd.backward([g1, bt, kfg], [dg1, dbt, dk])
    for k in ("w0", "b0", "w2", "b2", "mask", "scale"):
        assert rel(pa[k].grad.cpu(), pb[k].grad.cpu()) < 5e-3, k


def test_pool_fwd_bwd():
    from spff_b200 import ops
    n, c, d, h, w = 2, 64, 5, 8, 12
    torch.manual_seed(0)
    x = torch.randn(n, c, d, h, w, device="cuda")
    xb = pm(x)
    gamma = torch.ones(c, device="cuda"); beta = torch.zeros(c, device="cuda")
    coef = _coef(xb, c, gamma, beta)
    y = torch.empty_like(xb)
    yp = torch.empty(n, d, h // 2, w // 2, c, dtype=torch.bfloat16, device="cuda")
    ops.norm_act_affine_apply(xb, coef, None, None, y, yp, c, 0.01)
    y2 = torch.empty_like(xb)
    ops.norm_act_apply(xb, coef, y2, c, 0.01)
    assert torch.equal(y, y2)
    ref_pool = F.max_pool3d(ncdhw(y), (1, 2, 2))
    assert torch.equal(ncdhw(yp), ref_pool)
    # backward: scatter to arg-max + skip add
    yr = ncdhw(y).requires_grad_()
    dpool = torch.randn(n, c, d, h // 2, w // 2, device="cuda")
    dpb = pm(dpool)
    F.max_pool3d(yr, (1, 2, 2)).backward(ncdhw(dpb))
    skip = torch.randn(n, c, d, h, w, device="cuda")
    sb = pm(skip)
    ref = ncdhw(sb) + yr.grad
    ops.maxpool_bwd_add(dpb, y, sb, c, True)
    assert rel(ncdhw(sb), ref) < 5e-3
    sb2 = torch.full_like(sb, float("nan"))
    ops.maxpool_bwd_add(dpb, y, sb2, c, False)
    assert rel(ncdhw(sb2), yr.grad) < 1e-6
    # the arg-max codes written by the forward give the same backward without reading y (ties included: y is bf16)
    codes = torch.full((n, d, h // 2, w // 2, c), 255, dtype=torch.uint8, device="cuda")
    y3, yp3 = torch.empty_like(xb), torch.empty_like(yp)
    ops.norm_act_affine_apply(xb, coef, None, None, y3, yp3, c, 0.01, pool_argmax=codes)
    assert torch.equal(y3, y) and torch.equal(yp3, yp) and int(codes.max()) <= 3
    sb3 = torch.full_like(sb, float("nan"))
    ops.maxpool_bwd_add_argmax(dpb, codes, sb3, c, False)
    assert torch.equal(sb3, sb2)
    sb4 = pm(skip)
    ops.maxpool_bwd_add_argmax(dpb, codes, sb4, c, True)
    assert torch.equal(sb4, sb)


@pytest.mark.parametrize("n,d,h,w", [(2, 5, 16, 16), (1, 5, 7, 9), (1, 3, 4, 4), (1, 5, 8, 136), (2, 5, 128, 128)])
def test_stem(n, d, h, w):
    from spff_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(n, 1, d, h, w, device="cuda")
    wt = torch.randn(32, 1, 3, 3, 3, device="cuda") * 0.2
    y = torch.empty(n, d, h, w, 32, dtype=torch.bfloat16, device="cuda")
    ops.conv3d_stem_fwd(x, wt, y, 32)
    ref = F.conv3d(x, wt, padding=1)
    assert rel(ncdhw(y), ref) < 5e-3
    dy = torch.randn(n, 32, d, h, w, device="cuda")
    dyb = pm(dy)
    dw = torch.full((32, 1, 3, 3, 3), float("nan"), device="cuda")
    ops.conv3d_stem_wgrad(x, dyb, 32, dw, 0.0)
    # the tensor-core weight gradient reads x rounded to bf16: exact against that, 2^-9-close to fp32 x
    xq = x.to(torch.bfloat16).double()
    refq = torch.nn.grad.conv3d_weight(xq, (32, 1, 3, 3, 3), ncdhw(dyb).double(), padding=1).float()
    refw = torch.nn.grad.conv3d_weight(x.double(), (32, 1, 3, 3, 3), ncdhw(dyb).double(), padding=1).float()
    print("stem wgrad rel vs bf16-x", rel(dw, refq), "vs fp32-x", rel(dw, refw))
    assert rel(dw, refq) < 1e-4
    assert rel(dw, refw) < 4e-3
    ops.conv3d_stem_wgrad(x, dyb, 32, dw, 1.0)
    assert rel(dw, 2 * refq) < 1e-4
    # forward with the statistics taken in the epilogue: same output bits, partial sums = fp32 sums of the accumulators
    from spff_b200._lib import Shape
    slots = ops.conv3d_stem_stat_slots(Shape(n, d, h, w))
    partial = torch.full((n, slots, 2, 32), float("nan"), device="cuda")
    y2 = torch.full_like(y, float("nan"))
    ops.conv3d_stem_fwd_stats(x, wt, y2, 32, partial)
    # default: the tensor-core kernel on bf16 hi + lo operand splits (three products, fp32 accumulation) - the same
    # bf16 output as the fp32 FMA kernel except where ~2^-16 moves a value across a rounding boundary (one ulp there)
    assert rel(ncdhw(y2), ref) < 5e-3
    differs = y2 != y
    assert float(differs.float().mean()) < 2e-2
    ulp = 2.0 ** -7 * torch.maximum(y.float().abs(), y2.float().abs()) + 1e-4     # + the split error where terms cancel
    assert bool(((y2.float() - y.float()).abs() <= ulp).all())
    from spff_b200 import _lib
    _lib.lib.spff_debug_set(8, 1)                        # the fp32 FMA kernel with the same epilogue: same bits
    try:
        y3 = torch.full_like(y, float("nan"))
        partial3 = torch.full_like(partial, float("nan"))
        ops.conv3d_stem_fwd_stats(x, wt, y3, 32, partial3)
    finally:
        _lib.lib.spff_debug_set(8, 0)
    assert torch.equal(y3, y)
    assert rel(partial3.double().sum(1).float(), partial.double().sum(1).float()) < 1e-4
    tot = partial.double().sum(1)                       # [n, 2, 32]
    assert float((tot[:, 0].float() - ref.sum(dim=(2, 3, 4))).abs().max()) < 1e-3 * float(ref.abs().sum(dim=(2, 3, 4)).max())
    assert rel(tot[:, 1].float(), (ref * ref).sum(dim=(2, 3, 4))) < 1e-4
    coef = torch.empty(n, 32, 4, device="cuda")
    gamma, beta = torch.rand(32, device="cuda") + 0.5, torch.randn(32, device="cuda")
    ops.in_coeffs_from_partials(partial, slots, gamma, beta, 1e-5, n, 32, d * h * w, coef)
    assert torch.allclose(coef[..., 2], ref.mean(dim=(2, 3, 4)), atol=1e-4)
    assert torch.allclose(coef[..., 3], 1.0 / torch.sqrt(ref.var(dim=(2, 3, 4), unbiased=False) + 1e-5), rtol=1e-3)


@pytest.mark.parametrize("n,d,h,w,cin,cout", [(2, 5, 8, 8, 64, 32), (1, 5, 4, 4, 128, 64), (1, 5, 2, 2, 256, 128),
                                              (1, 3, 6, 10, 64, 32), (1, 5, 16, 16, 64, 32)])
def test_convt(n, d, h, w, cin, cout):
    from spff_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(n, cin, d, h, w, device="cuda")
    wt = torch.randn(cin, cout, 1, 2, 2, device="cuda") * (1.0 / cin ** 0.5)
    b = torch.randn(cout, device="cuda") * 0.1
    xb = pm(x)
    wf, wd = ops.pack_convt_weight(wt)
    # forward into the first half of a concat buffer
    cat = torch.zeros(n, d, 2 * h, 2 * w, 2 * cout, dtype=torch.bfloat16, device="cuda")
    ops.convt_k122_fwd(xb, cin, wf, b, cat[..., :cout], cout)
    wr = wt.to(torch.bfloat16).float()
    ref = F.conv_transpose3d(ncdhw(xb), wr, b, stride=(1, 2, 2))
    assert rel(ncdhw(cat[..., :cout]), ref) < 5e-3
    assert (cat[..., cout:] == 0).all()
    # dgrad from a strided dy
    dy = torch.randn(n, cout, d, 2 * h, 2 * w, device="cuda")
    dyb = pm(dy, ld=2 * cout)
    dx = torch.full((n, d, h, w, cin), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.convt_k122_dgrad(dyb, cout, wd, dx, cin)
    refdx = F.conv3d(ncdhw(dyb), wr, stride=(1, 2, 2))
    assert rel(ncdhw(dx), refdx) < 5e-3
    dw = torch.full((cin, cout, 1, 2, 2), float("nan"), device="cuda")
    ops.convt_k122_wgrad(xb, cin, dyb, cout, dw, 0.0)
    xr = ncdhw(xb).double().requires_grad_(False)
    wq = wt.double().clone().requires_grad_()
    F.conv_transpose3d(xr, wq, None, stride=(1, 2, 2)).backward(ncdhw(dyb).double())
    assert rel(dw, wq.grad.float()) < 1e-4


def test_head_and_loss():
    from spff_b200 import ops
    n, d, h, w, k = 2, 5, 12, 12, 13
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    x = torch.randn(n, 32, d, h, w, device="cuda")
    xb = pm(x)
    wt = torch.randn(k, 32, 1, 1, 1, device="cuda") * 0.3
    b = torch.randn(k, device="cuda") * 0.1
    logits = torch.empty(n, k, d, h, w, device="cuda")
    ops.head_fwd(xb, wt, b, logits)
    xr = ncdhw(xb).requires_grad_()
    wr = wt.clone().requires_grad_(); br = b.clone().requires_grad_()
    ref = F.conv3d(xr, wr, br)
    assert rel(logits, ref) < 1e-5
    lab8 = torch.empty(n, d, h, w, dtype=torch.uint8, device="cuda")
    ops.head_argmax(xb, wt, b, lab8)
    assert (lab8.long() == ref.argmax(1)).float().mean() > 0.9999
    labels = torch.randint(0, k, (n, d, h, w), device="cuda")
    labels[torch.rand(n, d, h, w, device="cuda") < 0.05] = 255
    acc = torch.zeros(1, dtype=torch.float64, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    conf = torch.zeros(k, k, dtype=torch.int64, device="cuda")
    ops.ce_confusion(logits, labels, 255, acc, cnt, conf)
    ce_ref = F.cross_entropy(ref, labels, ignore_index=255)
    assert int(cnt) == int((labels != 255).sum())
    assert abs(float(acc) / int(cnt) - float(ce_ref)) < 1e-5
    from oracle import spff_oracle as O
    cm = O.confusion(ref.argmax(1).cpu(), labels.cpu(), k, 255)
    assert (conf.cpu().numpy() == cm).all()
    # uint8 labels give the same tally
    conf2 = torch.zeros_like(conf); acc2 = torch.zeros_like(acc); cnt2 = torch.zeros_like(cnt)
    ops.ce_confusion(logits, labels.to(torch.uint8), 255, acc2, cnt2, conf2)
    assert torch.equal(conf, conf2) and int(cnt2) == int(cnt)
    dlog = torch.empty_like(logits)
    ops.ce_grad(logits, labels, 255, cnt, None, dlog)
    ce_ref.backward()
    lg = logits.clone().requires_grad_()
    F.cross_entropy(lg, labels, ignore_index=255).backward()
    assert rel(dlog, lg.grad) < 1e-4
    dx = torch.empty_like(xb)
    dw = torch.zeros(k, 32, device="cuda"); db = torch.zeros(k, device="cuda")
    ops.head_bwd(dlog, xb, wt, dx, dw, db, 0.0)
    assert rel(ncdhw(dx), xr.grad) < 5e-3
    assert rel(dw, wr.grad.reshape(k, 32)) < 1e-4
    assert rel(db, br.grad) < 1e-4


@pytest.mark.parametrize("n,d,h,w,dtype", [(2, 5, 12, 12, torch.int64), (1, 5, 7, 9, torch.uint8), (3, 5, 32, 32, torch.int64)])
def test_head_loss_fused(n, d, h, w, dtype):
    """Fused head + CE + confusion + backward == F.conv3d -> F.cross_entropy -> autograd (fp32)."""
    from spff_b200 import ops
    from oracle import spff_oracle as O
    k = 13
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1)
    xb = pm(torch.randn(n, 32, d, h, w, device="cuda"), ld=64)      # a channel slice of a wider buffer
    wt = (torch.randn(k, 32, 1, 1, 1, device="cuda") * 0.5).requires_grad_()
    b = (torch.randn(k, device="cuda") * 0.1).requires_grad_()
    labels = torch.randint(0, k, (n, d, h, w), device="cuda")
    labels[torch.rand(n, d, h, w, device="cuda") < 0.1] = 255
    xr = ncdhw(xb).requires_grad_()
    logits = F.conv3d(xr, wt, b)
    ce = F.cross_entropy(logits, labels, ignore_index=255)
    (ce * 0.5).backward()
    acc = torch.zeros(1, dtype=torch.float64, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    conf = torch.zeros(k, k, dtype=torch.int64, device="cuda")
    n_valid = (labels != 255).sum().reshape(1)
    gscale = torch.tensor([0.5], device="cuda")
    dx = torch.full((n, d, h, w, 64), float("nan"), dtype=torch.bfloat16, device="cuda")[..., :32]
    dw = torch.ones(k, 32, device="cuda"); db = torch.ones(k, device="cuda")
    ops.head_loss_fused(xb, wt.detach(), b.detach(), labels.to(dtype), 255, n_valid, gscale, acc, cnt, conf, dx, dw, db, 1.0)
    assert int(cnt) == int(n_valid)
    assert abs(float(acc) / int(cnt) - float(ce)) < 1e-5
    assert (conf.cpu().numpy() == O.confusion(logits.argmax(1).cpu(), labels.cpu(), k, 255)).all()
    assert rel(ncdhw(dx), xr.grad) < 5e-3
    assert rel(dw - 1, wt.grad.reshape(k, 32)) < 1e-4
    assert rel(db - 1, b.grad) < 1e-4
    # dx optional, beta = 0 overwrites
    dw2 = torch.full((k, 32), float("nan"), device="cuda"); db2 = torch.full((k,), float("nan"), device="cuda")
    acc.zero_(); cnt.zero_(); conf.zero_()
    ops.head_loss_fused(xb, wt.detach(), b.detach(), labels.to(dtype), 255, n_valid, gscale, acc, cnt, conf, None, dw2, db2, 0.0)
    assert rel(dw2, wt.grad.reshape(k, 32)) < 1e-4 and rel(db2, b.grad) < 1e-4


def test_adam():
    from spff_b200 import ops
    torch.manual_seed(0)
    p = torch.randn(10007, device="cuda")
    ref = p.clone().requires_grad_()
    opt = torch.optim.Adam([ref], lr=1e-3)
    m = torch.zeros_like(p); v = torch.zeros_like(p)
    for step in range(1, 4):
        g = torch.randn_like(p)
        ref.grad = g.clone()
        opt.step()
        ops.adam_step(p, g, m, v, 1e-3, 0.9, 0.999, 1e-8, step)
    assert torch.allclose(p, ref.detach(), atol=1e-6, rtol=1e-5)


def test_loss_tally_cache_is_keyed_by_tensor_identity():
    """Two different logits tensors that happen to reuse the same device block must not share a tally."""
    from innovative3D import helpers as H
    torch.manual_seed(0)
    lab = torch.randint(0, 13, (1, 5, 8, 8), device="cuda")
    vals = []
    for scale in (1.0, 5.0):
        logits = torch.randn(1, 13, 5, 8, 8, device="cuda") * scale      # freed each iteration -> same block next time
        vals.append((float(H.ce_plus_macro_dice_loss(logits, lab, 13)), float(F.cross_entropy(logits, lab)),
                     H.per_class_metrics_3d(logits, lab, 13, ignore_index=255)[3]))
        del logits
    assert abs(vals[0][0] - vals[1][0]) > 0.1
    for loss, ce, _ in vals:
        assert loss >= ce - 1e-5 and loss <= ce + 0.5 + 1e-5
