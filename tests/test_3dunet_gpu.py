"""GPU parity of the "3DUNet" control (Cicek 3D U-Net + depth adapter): the kernels only this variant uses
((2,2,2) transposed conv / pool, BatchNorm coefficients, depth resample, SGD) against PyTorch fp32 formulas, and
the whole network through innovative3D.models -> C ABI against the reference fixtures (tests/golden/cicek*.npz)
and the CPU oracle (oracle/cicek_oracle.py).
Tolerances: bf16-stored tensors rel-L2 <= 5e-3 per kernel; whole network logits / gradients 2e-2 relative
(BASELINE.json north_star) on trained weights, looser on the untrained name-seeded fixture weights where stated."""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(glob.glob(os.path.join(GOLD, "cicek[0-9]*.npz")))


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def pm(x, ld=None):  # NCDHW fp32 -> position-major bf16
    n, c, d, h, w = x.shape
    ld = ld or c
    buf = torch.zeros(n, d, h, w, ld, dtype=torch.bfloat16, device=x.device)
    buf[..., :c] = x.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
    return buf[..., :c] if ld != c else buf


def ncdhw(buf):
    return buf.permute(0, 4, 1, 2, 3).float()


# ------------------------------------------------------------------------------------------------
# kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,cin,cout,d,h,w", [(2, 64, 32, 8, 8, 8), (1, 128, 64, 4, 8, 12), (2, 512, 256, 1, 2, 2),
                                              (1, 256, 128, 2, 4, 4), (1, 512, 256, 1, 1, 1)])
def test_convt_k222(n, cin, cout, d, h, w):
    from spff_b200 import ops
    torch.manual_seed(1)
    x = torch.randn(n, cin, d, h, w, device="cuda")
    wt = torch.randn(cin, cout, 2, 2, 2, device="cuda") / cin ** 0.5
    bias = torch.randn(cout, device="cuda") * 0.1
    xb = pm(x)
    wf, wd = ops.pack_convt_weight_k222(wt)
    # forward into the low half of a 2*cout-channel buffer (the skip-concat layout)
    cat = torch.zeros(n, 2 * d, 2 * h, 2 * w, 2 * cout, dtype=torch.bfloat16, device="cuda")
    ops.convt_k222_fwd(xb, cin, wf, bias, cat[..., :cout], cout)
    wq = wt.to(torch.bfloat16).float()
    ref = F.conv_transpose3d(ncdhw(xb), wq, bias, stride=2)
    assert rel(ncdhw(cat[..., :cout]), ref) < 5e-3
    assert float(cat[..., cout:].abs().max()) == 0.0
    # dgrad / wgrad
    dy = torch.randn(n, cout, 2 * d, 2 * h, 2 * w, device="cuda")
    dcat = torch.zeros_like(cat)
    dcat[..., :cout] = dy.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
    dyv = dcat[..., :cout]
    dx = torch.empty(n, d, h, w, cin, dtype=torch.bfloat16, device="cuda")
    ops.convt_k222_dgrad(dyv, cout, wd, dx, cin)
    ref_dx = F.conv3d(ncdhw(dyv), wq, None, stride=2)
    assert rel(ncdhw(dx), ref_dx) < 5e-3
    dw = torch.full((cin, cout, 2, 2, 2), 0.5, device="cuda")
    ops.convt_k222_wgrad(xb, cin, dyv, cout, dw, 1.0)
    xr = ncdhw(xb).requires_grad_(False)
    wr = wq.clone().requires_grad_(True)
    (F.conv_transpose3d(xr, wr, None, stride=2) * ncdhw(dyv)).sum().backward()
    assert rel(dw - 0.5, wr.grad) < 2e-3


@pytest.mark.parametrize("din,dout", [(5, 16), (16, 5)])
def test_depth_resample(din, dout):
    from spff_b200 import ops
    from spff_b200.cicek import depth_matrix
    torch.manual_seed(2)
    m = depth_matrix(din, dout).cuda()
    x = torch.randn(3, 1, din, 8, 12, device="cuda")
    y = torch.empty(3, 1, dout, 8, 12, device="cuda")
    ops.depth_resample(x.view(3, din, -1), y.view(3, dout, -1), m)
    ref = F.interpolate(x, size=(dout, 8, 12), mode="trilinear", align_corners=False)
    assert float((y - ref).abs().max()) < 1e-5
    xb = torch.randn(2, din, 4, 6, 32, device="cuda").to(torch.bfloat16)
    yb = torch.empty(2, dout, 4, 6, 32, dtype=torch.bfloat16, device="cuda")
    ops.depth_resample(xb, yb, m)
    refb = F.interpolate(ncdhw(xb), size=(dout, 4, 6), mode="trilinear", align_corners=False)
    assert rel(ncdhw(yb), refb) < 4e-3
    # backward = the transposed matrix
    g = torch.randn(2, dout, 4, 6, 32, device="cuda").to(torch.bfloat16)
    gx = torch.empty_like(xb)
    ops.depth_resample(g, gx, m.t().contiguous())
    xr = ncdhw(xb).requires_grad_(True)
    (F.interpolate(xr, size=(dout, 4, 6), mode="trilinear", align_corners=False) * ncdhw(g)).sum().backward()
    assert rel(ncdhw(gx), xr.grad) < 4e-3


@pytest.mark.parametrize("n,c,d,h,w", [(3, 32, 4, 8, 8), (2, 512, 2, 2, 2), (2, 64, 16, 16, 16)])
def test_batchnorm_relu_fwd_bwd(n, c, d, h, w):
    """conv-epilogue partials -> spff_bn_coeffs -> normalise + ReLU; backward through bn_bwd_coeffs."""
    from spff_b200 import ops
    from spff_b200._lib import Shape
    torch.manual_seed(3)
    xin = torch.randn(n, c, d, h, w, device="cuda")
    wt = torch.randn(c, c, 3, 3, 3, device="cuda") / (27 * c) ** 0.5
    gamma = (torch.rand(c, device="cuda") + 0.5).requires_grad_(True)
    beta = (torch.randn(c, device="cuda") * 0.1).requires_grad_(True)
    rm, rv = torch.randn(c, device="cuda") * 0.1, torch.rand(c, device="cuda") + 0.5
    xb = pm(xin)
    wf, _ = ops.pack_conv3_weight(wt)
    y = torch.empty_like(xb)
    slots = ops.conv3d_k3_stat_slots(Shape(n, d, h, w))
    partial = torch.empty(n, slots, 2, c, device="cuda")
    ops.conv3d_k3_fwd_stats(xb, c, wf, y, c, partial)
    coef = torch.empty(n, c, 4, device="cuda")
    rm_k, rv_k = rm.clone(), rv.clone()
    ops.bn_coeffs(gamma.detach(), beta.detach(), 1e-5, n, c, d * h * w, coef, partial=partial, slots=slots,
                  running_mean=rm_k, running_var=rv_k)
    out = torch.empty_like(y)
    ops.norm_act_affine_apply(y, coef, None, None, out, None, c, 0.0)
    # fp32 reference on the conv output the kernel stored (statistics are taken before the bf16 rounding: 1e-3)
    yr = ncdhw(y).requires_grad_(True)
    rm_r, rv_r = rm.clone(), rv.clone()
    ref = F.relu(F.batch_norm(yr, rm_r, rv_r, gamma, beta, True, 0.1, 1e-5))
    assert rel(ncdhw(out), ref) < 6e-3
    assert torch.allclose(rm_k, rm_r, atol=2e-3) and torch.allclose(rv_k, rv_r, rtol=5e-3, atol=1e-4)
    assert float((coef - coef[:1]).abs().max()) == 0.0      # broadcast to every sample
    # the fp64-statistics entry (the stem's path) gives the same coefficients
    stats = torch.zeros(n, c, 2, dtype=torch.float64, device="cuda")
    ops.in_stats(y, c, stats)
    coef2 = torch.empty_like(coef)
    ops.bn_coeffs(gamma.detach(), beta.detach(), 1e-5, n, c, d * h * w, coef2, stats=stats)
    assert torch.allclose(coef2[..., 3], coef[..., 3], rtol=5e-3)
    # eval mode
    coef3 = torch.empty_like(coef)
    ops.bn_coeffs(gamma.detach(), beta.detach(), 1e-5, n, c, d * h * w, coef3, running_mean=rm, running_var=rv, eval_mode=True)
    oute = torch.empty_like(y)
    ops.norm_act_affine_apply(y, coef3, None, None, oute, None, c, 0.0)
    refe = F.relu(F.batch_norm(ncdhw(y), rm.clone(), rv.clone(), gamma, beta, False, 0.1, 1e-5))
    assert rel(ncdhw(oute), refe) < 5e-3
    # backward
    dout = torch.randn(n, c, d, h, w, device="cuda")
    # an element whose pre-activation is within rounding of 0 may take either ReLU branch (x*A+B in the kernel vs
    # (x-mean)*rstd*gamma+beta in ATen): give those no incoming gradient
    with torch.no_grad():
        zpre = F.batch_norm(ncdhw(y), rm.clone(), rv.clone(), gamma, beta, True, 0.0, 1e-5)
        dout[zpre.abs() < 1e-4] = 0.0
    db = pm(dout)
    ref.backward(ncdhw(db))
    R = torch.zeros(n, d, c, 6, device="cuda")
    ops.norm_act_bwd_reduce(db, y, coef, R, c, 0.0, plain=True, fixed_order=True)
    bcoef = torch.empty(n, c, 4, device="cuda")
    dg, dbt = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    ops.bn_bwd_coeffs(R, coef, gamma.detach(), c, Shape(n, d, h, w), bcoef, dg, dbt)
    dx = torch.empty_like(y)
    ops.norm_act_bwd_apply(db, y, coef, bcoef, None, None, dx, c, 0.0)
    assert rel(ncdhw(dx), yr.grad) < 1e-2
    assert rel(dg, gamma.grad) < 5e-3 and rel(dbt, beta.grad) < 5e-3


@pytest.mark.parametrize("n,c,d,h,w", [(2, 32, 4, 8, 8), (1, 256, 2, 2, 6), (3, 64, 16, 16, 16)])
def test_maxpool222(n, c, d, h, w):
    from spff_b200 import ops
    torch.manual_seed(4)
    x = torch.randn(n, c, d, h, w, device="cuda")
    x[0, :, :2, :2, :2] = 1.5          # ties: the first maximum in (d,h,w) order takes the gradient
    cat = torch.zeros(n, d, h, w, 2 * c, dtype=torch.bfloat16, device="cuda")
    cat[..., c:] = x.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
    yv = cat[..., c:]
    yp = torch.empty(n, d // 2, h // 2, w // 2, c, dtype=torch.bfloat16, device="cuda")
    ops.maxpool222_fwd(yv, yp, c)
    xr = ncdhw(yv).requires_grad_(True)
    ref = F.max_pool3d(xr, 2)
    assert torch.equal(ncdhw(yp), ref)
    g = torch.randn(n, c, d // 2, h // 2, w // 2, device="cuda")
    gb = pm(g)
    ref.backward(ncdhw(gb))
    prior = torch.randn(n, c, d, h, w, device="cuda")
    dcat = torch.zeros_like(cat)
    dcat[..., c:] = prior.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
    want = (ncdhw(dcat[..., c:]) + xr.grad).to(torch.bfloat16).float()
    ops.maxpool222_bwd_add(gb, yv, dcat[..., c:], c, True)
    assert torch.equal(ncdhw(dcat[..., c:]), want)
    ops.maxpool222_bwd_add(gb, yv, dcat[..., c:], c, False)
    assert torch.equal(ncdhw(dcat[..., c:]), xr.grad)


@pytest.mark.parametrize("nesterov,wd", [(False, 0.0), (True, 1e-3)])
def test_sgd_matches_torch(nesterov, wd):
    from spff_b200 import ops
    torch.manual_seed(5)
    p = torch.randn(10007, device="cuda")
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.SGD([ref], lr=1e-2, momentum=0.99, nesterov=nesterov, weight_decay=wd)
    buf = torch.zeros_like(p)
    for step in range(4):
        g = torch.randn_like(p)
        ref.grad = g.clone()
        opt.step()
        ops.sgd_step(p, g * 2.0, buf, 1e-2, 0.99, wd, nesterov, step == 0, 0.5)
        assert torch.allclose(p, ref.detach(), rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------------
# whole network
# ------------------------------------------------------------------------------------------------
def build():
    from innovative3D import config as C
    return dict((v[0], v[1]) for v in C.VARIANTS)["3DUNet"]().cuda()


def load(lit, weights):
    lit.load_state_dict(weights, strict=True)
    lit.backbone.materialize()


def grad_error(G, ref_grads, prefix="backbone."):
    num = sum(float((G[n].cpu().double() - ref_grads[prefix + n].double()).pow(2).sum()) for n in G)
    den = sum(float(ref_grads[prefix + n].double().pow(2).sum()) for n in G)
    return (num / den) ** 0.5


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p) for p in CASES])
def test_against_reference_fixture(path):
    """logits / loss / running statistics / gradients / eval logits vs what the reference itself produced
    (name-seeded untrained weights: deep-layer bf16 noise -> 4e-2 on logits, whole-gradient 0.1; see the
    trained-weights test for the 2e-2 bound)."""
    from oracle import cicek_oracle as CO
    from oracle import spff_oracle as O
    z = np.load(path)
    b, h, w, ign, seed = [str(v) for v in z["case"]]
    b, h, w, ign, seed = int(b), int(h), int(w), float(ign), int(seed)
    lit = build()
    weights = CO.det_weights(seed=42)
    load(lit, weights)
    x, lab = O.phantom_batch(b, h, w, seed=seed, ignore_frac=ign)
    lit.train()
    logits = lit(x.cuda())
    ref = torch.from_numpy(z["logits"])
    assert logits.shape == ref.shape
    # 2 x 16 x 16: the bottleneck holds ONE position per sample, so its BatchNorm normalises over 2 values per
    # channel (xhat = +-1 unless the two are within sqrt(eps) of each other, where bf16 rounding of the inputs
    # decides) — a degenerate statistic; the cases with >= 4 bottleneck positions hold 4e-2
    positions = b * (h // 16) * (w // 16)
    tol = 4e-2 if positions >= 4 else 0.3
    assert rel(logits, ref) < tol
    loss = lit._weighted_softmax_ce(logits, lab.cuda())
    assert abs(float(loss) - float(z["loss"])) < (3e-2 if positions >= 4 else 0.1)
    loss.backward()
    sd = lit.state_dict()
    for name in [str(n) for n in z["buf_names"]]:
        got, want = sd[name].cpu(), torch.from_numpy(z["b|" + name])
        assert float((got - want).norm() / (want.norm() + 1e-6)) < (3e-2 if positions >= 4 else 0.2), name
    assert int(sd["backbone.enc1.1.num_batches_tracked"]) == int(z["nbt"])
    names = [str(n) for n in z["grad_names"]]
    params = dict(lit.named_parameters())
    num = den = 0.0
    for name, gn in zip(names, z["grad_norms"]):
        g = params[name].grad.detach().double().reshape(-1).cpu()
        step = max(1, g.numel() // 512)
        s_ref = torch.from_numpy(z["g|" + name]).double()
        num += float((g[::step][:512] - s_ref).pow(2).sum())
        den += float(s_ref.pow(2).sum())
    if positions < 4:      # degenerate batch statistics (above): the gradient is not comparable at bf16
        return
    # untrained name-seeded weights: ReLU masks that flip under bf16 rounding compound through the 18 conv layers
    # of the backward pass; PyTorch's own CPU bf16 autocast of the oracle is 0.30-0.48 off on the encoder gradients
    # of these cases (oracle/probe_autocast_bf16_3dunet.py). The trained-weights test below holds the strict bound.
    assert (num / den) ** 0.5 < 0.6
    # eval mode: running statistics (reload: the training forward above moved them)
    load(lit, weights)
    lit.eval()
    with torch.no_grad():
        le = lit(x.cuda())
    assert rel(le, torch.from_numpy(z["logits_eval"])) < 4e-2    # running statistics: no degenerate batch statistic


def _trained(steps, b, h, w):
    """`steps` fused SGD steps on phantom batches (real margins, non-degenerate BatchNorm statistics)."""
    from oracle import cicek_oracle as CO
    from oracle import spff_oracle as O
    lit = build()
    load(lit, CO.det_weights(seed=42))
    lit.train()
    lit.hparams["lr"] = 2e-3
    for i in range(steps):
        x, lab = O.phantom_batch(b, h, w, seed=900 + i)
        lit.fit_step((x.cuda(), lab.cuda()))
    return lit


def _autocast_yardstick(weights, x, lab):
    """PyTorch's own CPU bf16 autocast of the oracle: (logits, grads) — the error floor of bf16 storage."""
    from oracle import cicek_oracle as CO
    q = {k: (v.detach().clone().requires_grad_(True) if not CO.is_buffer(k) else v.clone()) for k, v in weights.items()}
    with torch.autocast("cpu", dtype=torch.bfloat16):
        lg = CO.forward(q, x, True, None)
    CO.ce_loss(lg.float(), lab).backward()
    return lg.float().detach(), {k: v.grad for k, v in q.items() if not CO.is_buffer(k)}


def test_trained_weights_parity_with_oracle():
    """On briefly trained weights, against the CPU fp32 oracle: loss 1e-2, logits rel-L2 <= 2e-2, argmax agreement,
    macro Dice, BatchNorm running statistics, eval-mode logits, fused argmax. Gradients: the whole-gradient error
    must stay under 2e-2 or under 1.25x what PyTorch's own CPU bf16 autocast of the oracle shows on the same
    weights and inputs (ReLU + small-batch BatchNorm make this network's bf16 floor higher than SPFF-UNet's)."""
    from innovative3D import helpers as H
    from oracle import cicek_oracle as CO
    from oracle import spff_oracle as O
    lit = _trained(60, 4, 32, 32)
    weights = {k: v.detach().cpu().clone() for k, v in lit.state_dict().items()}
    x, lab = O.phantom_batch(4, 32, 32, seed=77, ignore_frac=0.01)
    ref_loss, ref_logits, ref_grads, ref_stats = CO.loss_and_grads(weights, x, lab)
    ac_logits, ac_grads = _autocast_yardstick(weights, x, lab)
    out = lit.fit_step((x.cuda(), lab.cuda()), optimize=False)
    assert abs(float(out["loss"]) - ref_loss) < 1e-2 * max(1.0, ref_loss)
    G = lit.fused_grads()
    err = grad_error(G, ref_grads)
    yard = grad_error({n: ac_grads["backbone." + n] for n in G}, ref_grads)
    print(f"whole-gradient rel-L2: B200 {err:.4f}, CPU bf16 autocast {yard:.4f}")
    assert err < max(2e-2, 1.25 * yard), (err, yard)
    bad = []
    for n in G:   # per layer: 3e-2, or twice the autocast yardstick of that layer
        r = ref_grads["backbone." + n]
        if float(r.norm()) > 1e-4:
            e, y = rel(G[n], r), rel(ac_grads["backbone." + n], r)
            if not e < max(3e-2, 2.0 * y):
                bad.append((n, round(e, 4), round(y, 4)))
    assert not bad, bad
    sd = lit.state_dict()
    for name, want in ref_stats.items():
        assert rel(sd[name], want) < 1e-2, name
    # forward (training-mode statistics, buffers restored first)
    load(lit, weights)
    lit.train()
    with torch.no_grad():
        logits = lit(x.cuda())
    e_log, y_log = rel(logits, ref_logits), rel(ac_logits, ref_logits)
    print(f"logits rel-L2: B200 {e_log:.4f}, CPU bf16 autocast {y_log:.4f}")
    assert e_log < 2e-2
    agree = float((logits.argmax(1).cpu() == ref_logits.argmax(1)).float().mean())
    agree_ac = float((ac_logits.argmax(1) == ref_logits.argmax(1)).float().mean())
    print(f"argmax agreement: B200 {agree:.5f}, CPU bf16 autocast {agree_ac:.5f}")
    assert agree >= min(0.999, agree_ac - 2e-3), (agree, agree_ac)
    m_gpu = H.per_class_metrics_3d(logits, lab.cuda(), 13, ignore_index=255)
    m_ref = O.per_class_metrics_3d(ref_logits, lab, 13, ignore_index=255)
    assert abs(m_gpu[3] - m_ref[3]) < 5e-3
    # eval mode + fused argmax
    load(lit, weights)
    lit.eval()
    with torch.no_grad():
        le = lit(x.cuda())
        ref_e = CO.forward(weights, x, training=False)
        labels = lit.predict_labels(x.cuda())
    assert rel(le, ref_e) < 2e-2
    assert float((labels.cpu().long() == le.argmax(1).cpu()).float().mean()) > 0.999


def test_fused_step_equals_autograd_path_and_sgd():
    """fit_step (fused head + CE, SGD kernel) == training_step + backward + torch.optim.SGD.step on the same weights."""
    from oracle import cicek_oracle as CO
    from oracle import spff_oracle as O
    weights = CO.det_weights(seed=42)
    x, lab = O.phantom_batch(2, 32, 32, seed=5, ignore_frac=0.02)
    a, b = build(), build()
    load(a, weights)
    load(b, weights)
    a.train(); b.train()
    opt = b.configure_optimizers()
    out = a.fit_step((x.cuda(), lab.cuda()))
    loss = b.training_step((x.cuda(), lab.cuda()), 0)
    loss.backward()
    assert abs(float(out["loss"]) - float(loss.detach())) < 1e-4
    pb = dict(b.backbone.named_parameters())
    G = a.fused_grads()
    num = sum(float((G[n] - pb[n].grad).double().pow(2).sum()) for n in G)
    den = sum(float(pb[n].grad.double().pow(2).sum()) for n in G)
    assert (num / den) ** 0.5 < 2e-2     # bf16 head-input gradient (fused kernel) vs fp32 dlogits -> head_bwd
    opt.step()
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        if sa[k].dtype.is_floating_point:
            assert rel(sa[k], sb[k]) < 1e-3, k
        else:
            assert int(sa[k]) == int(sb[k]) == 1, k
    # a second fused step uses the momentum buffer and the re-packed weights
    out2 = a.fit_step((x.cuda(), lab.cuda()))
    assert torch.isfinite(out2["loss"]) and float(out2["loss"]) < float(out["loss"])
    assert int(a.state_dict()["backbone.bott.4.num_batches_tracked"]) == 2


def test_native_slice_size_step_runs():
    """One fused step at the native slice size [2,1,5,128,128] (16 planes inside): finite loss, every gradient finite."""
    lit = build()
    lit.train()
    torch.manual_seed(0)
    x = torch.randn(2, 1, 5, 128, 128, device="cuda")
    lab = torch.randint(0, 13, (2, 5, 128, 128), device="cuda")
    out = lit.fit_step((x, lab))
    assert torch.isfinite(out["loss"])
    assert all(bool(torch.isfinite(g).all()) for g in lit.fused_grads().values())


def test_errors():
    lit = build()
    with pytest.raises(ValueError):
        lit(torch.zeros(1, 1, 5, 24, 16, device="cuda"))     # H not a multiple of 16
    with pytest.raises(RuntimeError):
        lit.backbone.engine.infer(torch.zeros(1, 1, 5, 16, 16), 16)
    from innovative3D.models import LitCicek3DUNet_DepthAdapter_Published as L
    with pytest.raises(NotImplementedError):
        L(num_classes=13, class_weights=[1.0] * 13)
    with pytest.raises(NotImplementedError):
        L(num_classes=13, dice_weight=0.5)



def test_eval_mode_gradients_and_second_backward_are_loud():
    """The backward kernels implement training-mode BatchNorm: asking for gradients in eval() raises instead of
    returning the training-mode gradient; a second backward through one forward raises a clear error."""
    from innovative3D import config as C
    lit = dict((v[0], v[1]) for v in C.VARIANTS)["3DUNet"]().cuda()
    x = torch.randn(2, 1, 5, 16, 16, device="cuda")
    lit.eval()
    with pytest.raises(NotImplementedError, match="eval-mode BatchNorm"):
        lit(x)
    with torch.no_grad():
        assert lit(x).shape == (2, 13, 5, 16, 16)          # inference in eval mode is fine
    lit.train()
    out = lit(x)
    out.sum().backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="second backward"):
        out.sum().backward()
    spff = dict((v[0], v[1]) for v in C.VARIANTS)["SPFF-UNet"]().cuda()
    o2 = spff(x)
    o2.sum().backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="second backward"):
        o2.sum().backward()
