"""GPU parity tests: every kernel is called through the C-ABI (ctypes) and compared with the CPU
oracle (oracle/torch_ref.py, pinned to the real reference by test_oracle_golden.py).
Tolerances: forward activations 1e-4 relative (max-abs / max-abs; the north-star bar on the logits is
1e-3), gradients 2e-3 relative."""
import importlib

import numpy as np
import pytest
import torch

from oracle import torch_ref as R
from util import ckpt_params, golden, load_test_wav, rel_err

pytestmark = pytest.mark.gpu
FWD_TOL, GRAD_TOL = 1e-4, 2e-3


@pytest.fixture(scope="module")
def pkg():
    p = importlib.import_module("end-to-end-slu_b200")
    p._lib.load()
    return p


def dev(t):
    return t.cuda().contiguous()


@pytest.mark.parametrize("N,K", [(16, 16), (16, 128), (32, 64), (80, 80), (128, 128), (240, 64), (256, 128)])
def test_tcgen05_selftest(pkg, N, K):
    rs = np.random.RandomState(N * 1000 + K)
    A = rs.standard_normal((128, K)).astype(np.float32)
    Bm = rs.standard_normal((N, K)).astype(np.float32)
    C = torch.empty(128, N, device="cuda")
    Ad, Bd = dev(torch.from_numpy(A)), dev(torch.from_numpy(Bm))        # keep alive: raw pointers are passed
    pkg._lib.call("slu_tc_selftest", Ad.data_ptr(), Bd.data_ptr(), C.data_ptr(), N, K, pkg._lib.stream())
    torch.cuda.synchronize()
    ref = A.astype(np.float64) @ Bm.astype(np.float64).T
    err = np.abs(C.cpu().numpy() - ref).max() / np.abs(ref).max()
    assert err < 2e-5, err          # 3-pass bf16 split ~ 2^-16; plain bf16 would be ~4e-3


@pytest.mark.parametrize("N,K", [(16, 16), (16, 128), (32, 64), (64, 128), (256, 32)])
def test_tcgen05_selftest_a_in_tmem(pkg, N, K):
    rs = np.random.RandomState(N * 1000 + K + 1)
    A = rs.standard_normal((128, K)).astype(np.float32)
    Bm = rs.standard_normal((N, K)).astype(np.float32)
    C = torch.empty(128, N, device="cuda")
    Ad, Bd = dev(torch.from_numpy(A)), dev(torch.from_numpy(Bm))        # keep alive: raw pointers are passed
    pkg._lib.call("slu_tc_selftest_ts", Ad.data_ptr(), Bd.data_ptr(), C.data_ptr(), N, K, pkg._lib.stream())
    torch.cuda.synchronize()
    ref = A.astype(np.float64) @ Bm.astype(np.float64).T
    err = np.abs(C.cpu().numpy() - ref).max() / np.abs(ref).max()
    assert err < 2e-5, err


def test_sinc_filters_fwd_bwd(pkg):
    for src in ("mel", "ckpt"):
        if src == "mel":
            p = R.synthetic_params()
        else:
            p = ckpt_params()
        b1 = p[R.P + "phoneme_layers.0.filt_b1"].clone().requires_grad_(True)
        band = p[R.P + "phoneme_layers.0.filt_band"].clone().requires_grad_(True)
        W_ref = R.sinc_filters(b1, band)
        W = pkg.ops.sinc_filters(b1.cuda(), band.cuda())
        assert rel_err(W.cpu(), W_ref) < 1e-5
        rs = np.random.RandomState(3)
        dW = torch.from_numpy(rs.standard_normal((80, 401)).astype(np.float32))
        W_ref.backward(dW)
        d_b1 = torch.empty(80, device="cuda", dtype=torch.float64); d_band = torch.empty_like(d_b1)
        b1d, bandd, dWd = dev(b1.detach()), dev(band.detach()), dev(dW)    # keep alive: raw pointers are passed
        pkg._lib.call("slu_sinc_filters_bwd", b1d.data_ptr(), bandd.data_ptr(), dWd.data_ptr(), d_b1.data_ptr(),
                      d_band.data_ptr(), pkg._lib.stream())
        e1, e2 = rel_err(d_b1.cpu(), b1.grad), rel_err(d_band.cpu(), band.grad)
        assert e1 < 1e-3 and e2 < 1e-3, (src, e1, e2)


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("B,T", [(2, 8000), (3, 12345), (2, 1234)])
def test_sincconv_filter_gradient_direct(pkg, impl, B, T):
    """dL/dW of the fused conv+abs+pool against autograd over the same bank (isolates the conv backward kernel)."""
    rs = np.random.RandomState(T)
    x = torch.from_numpy((0.1 * rs.standard_normal((B, T))).astype(np.float32))
    W0 = torch.from_numpy(rs.standard_normal((80, 401)).astype(np.float32))
    W = (W0 + W0.flip(1)).requires_grad_(True)          # sinc banks are symmetric; the simt kernel relies on it
    ref = torch.nn.functional.max_pool1d(torch.abs(torch.nn.functional.conv1d(x.unsqueeze(1), W.unsqueeze(1), stride=80, padding=200)),
                                         2, ceil_mode=True).transpose(1, 2)
    gy = torch.from_numpy(rs.standard_normal(tuple(ref.shape)).astype(np.float32))
    ref.backward(gy)
    L1 = ref.shape[1]
    xd, Wd, gyd = x.cuda(), W.detach().cuda(), gy.cuda().contiguous()
    out = torch.empty(B, L1, 80, device="cuda"); route = torch.empty(B, L1, 80, device="cuda", dtype=torch.uint8)
    dW = torch.zeros(80, 401, device="cuda")
    if impl == "tc":
        img = torch.empty(2 * 6 * 80 * 96, device="cuda", dtype=torch.bfloat16)
        pkg._lib.call("slu_sincconv_fwd_tc", xd.data_ptr(), Wd.data_ptr(), B, T, out.data_ptr(), route.data_ptr(), img.data_ptr(), pkg._lib.stream())
        pkg._lib.call("slu_sincconv_bwd_tc", xd.data_ptr(), gyd.data_ptr(), route.data_ptr(), B, T, dW.data_ptr(), pkg._lib.stream())
    else:
        pkg._lib.call("slu_sincconv_fwd_simt", xd.data_ptr(), Wd.data_ptr(), B, T, out.data_ptr(), route.data_ptr(), pkg._lib.stream())
        pkg._lib.call("slu_sincconv_bwd_simt", xd.data_ptr(), gyd.data_ptr(), route.data_ptr(), B, T, dW.data_ptr(), pkg._lib.stream())
    assert rel_err(out.cpu(), ref.detach()) < FWD_TOL
    err = rel_err(dW.cpu(), W.grad)
    assert err < 1e-4, err


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("B,T", [(1, 57585), (3, 8000), (2, 1234), (2, 81), (1, 1), (5, 64000), (3, 12345)])
def test_sinc_frontend_fwd_bwd(pkg, monkeypatch, impl, B, T):
    monkeypatch.setattr(pkg.ops, "SINC_IMPL", impl)
    p = ckpt_params() if T == 57585 else R.synthetic_params()
    x = load_test_wav() if T == 57585 else R.synthetic_batch(B, T, seed=T)[0]
    b1 = p[R.P + "phoneme_layers.0.filt_b1"].clone().requires_grad_(True)
    band = p[R.P + "phoneme_layers.0.filt_band"].clone().requires_grad_(True)
    ref = R.sinc_frontend(x, b1, band).transpose(1, 2)                      # NLC
    b1g = b1.detach().cuda().requires_grad_(True); bandg = band.detach().cuda().requires_grad_(True)
    out = pkg.ops.SincFrontend.apply(x.cuda(), b1g, bandg)
    assert out.shape == ref.shape
    assert rel_err(out.detach().cpu(), ref) < FWD_TOL
    rs = np.random.RandomState(1)
    gy = torch.from_numpy(rs.standard_normal(tuple(ref.shape)).astype(np.float32))
    ref.backward(gy)
    out.backward(gy.cuda())
    assert b1g.grad.dtype == torch.float64
    scale = max(b1.grad.abs().max().item(), band.grad.abs().max().item())
    for got, ref_g in ((b1g.grad.cpu(), b1.grad), (bandg.grad.cpu(), band.grad)):
        # abs / max-pool ROUTING is discontinuous: the tcgen05 forward differs from fp32 by ~1e-5, which flips the winner of a
        # fraction ~2e-5 of the frame pairs; each flip moves one of N random-signed terms, so the cut-off gradients move by
        # ~sqrt(2e-5) = 4.5e-3 relative whatever N is (measured 2e-3..7e-3).  test_sinc_cutoff_gradients_given_the_routing pins the
        # kernel itself to 1e-4; here the bound is the routing noise.
        tol = (1e-2 if impl == "tc" else GRAD_TOL) * scale + 1e-4
        assert (got - ref_g).abs().max().item() < tol, ((got - ref_g).abs().max().item(), scale)


@pytest.mark.parametrize("B,T", [(3, 12345), (2, 64000), (2, 1234)])
def test_sinc_cutoff_gradients_given_the_routing(pkg, B, T):
    """slu_sinc_filters_jac + slu_sincconv_bwd_jac_tc against autograd of the SAME routed objective: with the abs / max-pool
    routing fixed to the route bits the CUDA forward produced, dL/d(filt_b1, filt_band) is a smooth function and the kernels
    must agree with fp32 autograd to 1e-4 (this isolates them from the forward's routing flips)."""
    p = R.synthetic_params()
    x = R.synthetic_batch(B, T, seed=T)[0]
    b1 = p[R.P + "phoneme_layers.0.filt_b1"].clone().requires_grad_(True)
    band = p[R.P + "phoneme_layers.0.filt_band"].clone().requires_grad_(True)
    L0 = (T - 1) // 80 + 1
    L1 = (L0 + 1) // 2
    rs = np.random.RandomState(2)
    gy = torch.from_numpy(rs.standard_normal((B, L1, 80)).astype(np.float32))
    xd, gyd = x.cuda(), gy.cuda()
    b1d, bandd = b1.detach().cuda(), band.detach().cuda()
    W = pkg.ops.sinc_filters(b1d, bandd)
    out = torch.empty(B, L1, 80, device="cuda"); route = torch.empty(B, L1, 80, device="cuda", dtype=torch.uint8)
    img = torch.empty(2 * 6 * 160 * 96, device="cuda", dtype=torch.bfloat16)
    st = pkg._lib.stream()
    pkg._lib.call("slu_sincconv_fwd_tc", xd.data_ptr(), W.data_ptr(), B, T, out.data_ptr(), route.data_ptr(), img.data_ptr(), st)
    J = torch.empty(2, 80, 401, device="cuda")
    d = torch.zeros(160, device="cuda", dtype=torch.float64)
    pkg._lib.call("slu_sinc_filters_jac", b1d.data_ptr(), bandd.data_ptr(), J.data_ptr(), st)
    pkg._lib.call("slu_sincconv_bwd_jac_tc", xd.data_ptr(), gyd.data_ptr(), route.data_ptr(), J.data_ptr(), B, T, d.data_ptr(),
                  img.data_ptr(), st)
    # the routed objective on the CPU: frame 2j + sel of every pair gets +-gy (0 where the winner was exactly 0)
    rt = route.cpu().long()                                            # [B, L1, 80]
    sel, neg, zero = rt & 1, (rt >> 1) & 1, (rt >> 2) & 1
    G0 = torch.zeros(B, 2 * L1, 80)
    val = gy * (1 - 2 * neg).float() * (1 - zero).float()
    G0.scatter_(1, (2 * torch.arange(L1).view(1, L1, 1) + sel), val)
    G0 = G0[:, :L0].transpose(1, 2)                                    # [B, 80, L0]
    Wr = R.sinc_filters(b1, band)
    conv = torch.nn.functional.conv1d(x.unsqueeze(1), Wr.unsqueeze(1), stride=80, padding=200)
    (conv * G0).sum().backward()
    scale = max(b1.grad.abs().max().item(), band.grad.abs().max().item())
    dc = d.cpu()
    assert (dc[:80] - b1.grad).abs().max().item() < 1e-4 * scale, (dc[:80] - b1.grad).abs().max().item() / scale
    assert (dc[80:] - band.grad).abs().max().item() < 1e-4 * scale, (dc[80:] - band.grad).abs().max().item() / scale


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("B,T,I,ds,use_mask", [(3, 7, 60, 2, False), (17, 9, 256, 2, True), (40, 5, 60, 1, False), (4, 24, 256, 2, True), (5, 23, 256, 1, True),
                                               (1, 1, 256, 2, False), (9, 50, 60, 2, False), (2, 360, 60, 2, False),
                                               # B >= 592 / >= 1184 select the 8- and 16-rows-per-CTA instantiations
                                               (601, 6, 60, 2, True), (1187, 5, 256, 1, False)])
def test_bigru_fwd_bwd(pkg, monkeypatch, impl, B, T, I, ds, use_mask):
    monkeypatch.setattr(pkg.ops, "GRU_IMPL", impl)
    rs = np.random.RandomState(B * 100 + T)
    gru = torch.nn.GRU(I, 128, batch_first=True, bidirectional=True)
    with torch.no_grad():
        for q in gru.parameters():
            q.copy_(torch.from_numpy(rs.uniform(-0.15, 0.15, size=tuple(q.shape)).astype(np.float32)))
    x = torch.from_numpy(rs.standard_normal((B, T, I)).astype(np.float32)).requires_grad_(True)
    mask = torch.from_numpy((rs.uniform(size=(B, T, 256)) >= 0.5).astype(np.float32) * 2) if use_mask else None
    params = {"g." + k: v for k, v in gru.named_parameters()}
    y = R.bigru(x, params, "g")
    y = R.downsample(R.apply_dropout(y, mask), "avg" if ds == 2 else "none", ds)
    gy = torch.from_numpy(rs.standard_normal(tuple(y.shape)).astype(np.float32))
    y.backward(gy)
    g_ref = {k: v.grad.clone() for k, v in gru.named_parameters()}
    gx_ref = x.grad.clone()
    gru_c = torch.nn.GRU(I, 128, batch_first=True, bidirectional=True).cuda()
    gru_c.load_state_dict(gru.state_dict())
    xc = x.detach().cuda().requires_grad_(True)
    yc = pkg.ops.bigru(xc, gru_c, None if mask is None else mask.cuda(), ds)
    assert yc.shape == y.shape
    assert rel_err(yc.detach().cpu(), y.detach()) < FWD_TOL
    yc.backward(gy.cuda())
    assert rel_err(xc.grad.cpu(), gx_ref) < GRAD_TOL
    for k, v in gru_c.named_parameters():
        assert rel_err(v.grad.cpu(), g_ref[k]) < GRAD_TOL, k


@pytest.mark.parametrize("M,K,N", [(300, 60, 768), (257, 256, 768), (128, 256, 24), (1000, 768, 60), (130, 768, 256)])
def test_gemm_tc_linear_and_input_grad(pkg, monkeypatch, M, K, N):
    rs = np.random.RandomState(M + K + N)
    x = torch.from_numpy(rs.standard_normal((M, K)).astype(np.float32)).cuda()
    w = torch.from_numpy(rs.standard_normal((N, K)).astype(np.float32)).cuda()
    b = torch.from_numpy(rs.standard_normal(N).astype(np.float32)).cuda()
    ref = x.double() @ w.double().t() + b.double()
    assert rel_err(pkg.ops.linear_nt(x, w, b).cpu(), ref.cpu()) < 2e-5
    wk = torch.from_numpy(rs.standard_normal((K, N)).astype(np.float32)).cuda()
    assert rel_err(pkg.ops.matmul_nn(x, wk).cpu(), (x.double() @ wk.double()).cpu()) < 2e-5


@pytest.mark.parametrize("R,M,N", [(1000, 768, 60), (5000, 768, 256), (333, 60, 80), (4096, 256, 128)])
def test_gemm_tc_weight_grad_splitk(pkg, monkeypatch, R, M, N):
    rs = np.random.RandomState(R + M + N)
    g = torch.from_numpy(rs.standard_normal((R, M)).astype(np.float32)).cuda()
    x = torch.from_numpy(rs.standard_normal((R, N)).astype(np.float32)).cuda()
    ref = g.double().t() @ x.double()
    assert rel_err(pkg.ops.matmul_tn(g, x).cpu(), ref.cpu()) < 2e-5


@pytest.mark.parametrize("B,T,Cin,Cout", [(3, 37, 80, 60), (2, 1, 60, 60), (5, 400, 60, 60), (1, 130, 80, 60)])
def test_conv_block_tc_fwd_bwd(pkg, monkeypatch, B, T, Cin, Cout):
    rs = np.random.RandomState(B * 1000 + T)
    x = torch.from_numpy(rs.standard_normal((B, Cin, T)).astype(np.float32)).requires_grad_(True)
    w = torch.from_numpy(rs.uniform(-0.1, 0.1, (Cout, Cin, 5)).astype(np.float32)).requires_grad_(True)
    b = torch.from_numpy(rs.uniform(-0.1, 0.1, Cout).astype(np.float32)).requires_grad_(True)
    ref = R.conv_block(x, w, b).transpose(1, 2)
    gy = torch.from_numpy(rs.standard_normal(tuple(ref.shape)).astype(np.float32))
    ref.backward(gy)
    xc = x.detach().transpose(1, 2).contiguous().cuda().requires_grad_(True)
    wc = w.detach().cuda().requires_grad_(True); bc = b.detach().cuda().requires_grad_(True)
    out = pkg.ops.conv_block(xc, wc, bc, 0.2)
    assert rel_err(out.detach().cpu(), ref.detach()) < FWD_TOL
    out.backward(gy.cuda())
    assert rel_err(xc.grad.cpu(), x.grad.transpose(1, 2)) < GRAD_TOL
    assert rel_err(wc.grad.cpu(), w.grad) < GRAD_TOL and rel_err(bc.grad.cpu(), b.grad) < GRAD_TOL


@pytest.mark.parametrize("mode,ftol,gtol,fused", [("fp16", 2e-3, 1e-2, True), ("bf16x3-separate", FWD_TOL, GRAD_TOL, True),
                                                  ("bf16x3", FWD_TOL, GRAD_TOL, False)])
@pytest.mark.parametrize("B,T,I,ds", [(16, 40, 256, 2), (5, 23, 60, 1)])
def test_bigru_alternative_operand_formats(pkg, monkeypatch, B, T, I, ds, mode, ftol, gtol, fused):
    """Other operand formats of the recurrence: one fp16 pass (looser per-layer tolerance) and the un-stacked three-pass
    bf16 split (same tolerance as the default stacked form); and the launch-by-launch backward (the default is one C call)."""
    monkeypatch.setattr(pkg.ops, "GRU_IMPL", "tc")
    monkeypatch.setattr(pkg.ops, "FUSED_BWD", fused)
    pkg.ops.set_gru_precision(mode)
    try:
        rs = np.random.RandomState(B + T)
        gru = torch.nn.GRU(I, 128, batch_first=True, bidirectional=True)
        x = torch.from_numpy(rs.standard_normal((B, T, I)).astype(np.float32)).requires_grad_(True)
        y = R.downsample(R.bigru(x, {"g." + k: v for k, v in gru.named_parameters()}, "g"), "avg" if ds == 2 else "none", ds)
        gy = torch.from_numpy(rs.standard_normal(tuple(y.shape)).astype(np.float32))
        y.backward(gy)
        gru_c = torch.nn.GRU(I, 128, batch_first=True, bidirectional=True).cuda()
        gru_c.load_state_dict(gru.state_dict())
        xc = x.detach().cuda().requires_grad_(True)
        yc = pkg.ops.bigru(xc, gru_c, None, ds)
        assert rel_err(yc.detach().cpu(), y.detach()) < ftol
        yc.backward(gy.cuda())
        assert rel_err(xc.grad.cpu(), x.grad) < gtol
        for k, v in gru_c.named_parameters():
            assert rel_err(v.grad.cpu(), gru.get_parameter(k).grad) < gtol, k
    finally:
        pkg.ops.set_gru_precision("bf16x3")


@pytest.mark.parametrize("n,p", [(1 << 20, 0.5), (1000003, 0.25), (7, 0.5), (3, 0.0)])
def test_dropout_mask_kernel(pkg, n, p):
    """nn.Dropout's training-mode mask (models.py:246): values in {0, 1/(1-p)}, keep rate 1-p, reproducible per seed."""
    L = pkg._lib

    def draw(seed):
        buf = torch.full((n + 8,), -1.0, device="cuda")
        L.call("slu_dropout_mask", L.ptr(buf), n, float(p), seed, L.stream())
        torch.cuda.synchronize()
        assert (buf[n:] == -1).all()                      # nothing written past n
        return buf[:n].cpu()
    a, b, c = draw(1234), draw(1234), draw(1235)
    scale = 1.0 / (1.0 - p)
    assert torch.equal(a, b)
    assert ((a == 0) | ((a - scale).abs() < 1e-6)).all()
    if n > 1000:
        keep = (a != 0).float()
        sigma = (p * (1 - p) / n) ** 0.5
        assert abs(keep.mean().item() - (1 - p)) < 5 * sigma
        assert not torch.equal(a, c)
        k0 = keep - keep.mean()
        for lag in (1, 2, 4, 256):                        # no visible correlation inside / across Philox counters
            corr = (k0[:-lag] * k0[lag:]).mean().item() / (p * (1 - p))
            assert abs(corr) < 5 / n ** 0.5, (lag, corr)
    if p == 0.0:
        assert (a == 1).all()


@pytest.mark.parametrize("B,T,slots", [(5, 7, (6, 14, 4)), (300, 25, (6, 14, 4)), (3, 94, (3, 5)), (9, 33, (7,)), (2, 1, (2, 2, 2, 2, 120))])
def test_intent_head_fwd_bwd(pkg, B, T, slots):
    """Fused Linear + max-over-time + per-slot CE + accuracy (models.py:709, 112-123, 811-823) against torch ops in fp64."""
    rs = np.random.RandomState(B * 7 + T)
    C = sum(slots)
    feats = torch.from_numpy(rs.standard_normal((B, T, 256)).astype(np.float32))
    W = torch.from_numpy(rs.uniform(-0.1, 0.1, size=(C, 256)).astype(np.float32))
    bias = torch.from_numpy(rs.uniform(-0.1, 0.1, size=(C,)).astype(np.float32))
    y = torch.stack([torch.from_numpy(rs.randint(0, n, size=(B,))) for n in slots], 1)
    f64, w64, b64 = (t.double().requires_grad_(True) for t in (feats, W, bias))
    logits_ref = (f64 @ w64.t() + b64).max(dim=1)[0]
    loss_ref, start, ok = 0.0, 0, torch.ones(B, dtype=torch.bool)
    for s, n in enumerate(slots):
        loss_ref = loss_ref + torch.nn.functional.cross_entropy(logits_ref[:, start:start + n], y[:, s])
        ok &= logits_ref[:, start:start + n].max(1)[1] == y[:, s]
        start += n
    (loss_ref * 1.7).backward()
    fc, wc, bc = (t.cuda().requires_grad_(True) for t in (feats, W, bias))
    loss, acc, logits = pkg.ops.IntentHead.apply(fc, wc, bc, y.cuda(), slots)
    assert loss.dim() == 0 and acc.dim() == 0 and not acc.requires_grad
    (loss * 1.7).backward()
    assert rel_err(logits.cpu(), logits_ref.detach().float()) < 1e-5
    assert abs(loss.item() - loss_ref.item()) < 1e-5 * abs(loss_ref.item())
    assert abs(acc.item() - ok.float().mean().item()) < 1e-6
    assert rel_err(fc.grad.cpu(), f64.grad.float()) < 1e-4
    assert rel_err(wc.grad.cpu(), w64.grad.float()) < 1e-4
    assert rel_err(bc.grad.cpu(), b64.grad.float()) < 1e-4
    again = pkg.ops.IntentHead.apply(fc, wc, bc, y.cuda(), slots)[0]
    assert again.item() == loss.item()                                   # fixed-order batch reduction: bit-reproducible
    assert rel_err(pkg.ops.intent_head_logits(fc, wc, bc).cpu(), logits_ref.detach().float()) < 1e-5


@pytest.mark.parametrize("B,T,shift", [(3, 37, -1), (2, 16, 1), (5, 1, -1), (4, 50, 0)])
def test_wgrad_two_source_rows_and_frame_shift(pkg, B, T, shift):
    """slu_wgrad2_tc as BiGRU.backward uses it: dW_hh[d] = [dgx[:, :, d*384 : d*384+256] | dhn[:, :, d*128 : +128]]^T . y shifted
    by one frame inside each utterance (zero across utterance boundaries)."""
    rs = np.random.RandomState(B * 10 + T)
    dgx = torch.from_numpy(rs.standard_normal((B, T, 768)).astype(np.float32))
    dhn = torch.from_numpy(rs.standard_normal((B, T, 256)).astype(np.float32))
    y = torch.from_numpy(rs.standard_normal((B, T, 256)).astype(np.float32))
    for d in range(2):
        G = torch.cat([dgx[:, :, d * 384:d * 384 + 256], dhn[:, :, d * 128:(d + 1) * 128]], 2).double()      # [B,T,384]
        Xs = torch.zeros(B, T, 128, dtype=torch.float64)
        yy = y[:, :, d * 128:(d + 1) * 128].double()
        if shift == -1:
            Xs[:, 1:] = yy[:, :-1]
        elif shift == 1:
            Xs[:, :-1] = yy[:, 1:]
        else:
            Xs = yy
        ref = torch.einsum("btm,btn->mn", G, Xs).float()
        out = torch.zeros(2, 384, 128, device="cuda")
        pkg.ops.wgrad2_tc(dgx.cuda(), d * 384, 768, 256, dhn.cuda(), d * 128, 256, 384, y.cuda(), d * 128, 256, 128, B, T, out,
                          d * 384 * 128, 128, shift0=shift)
        torch.cuda.synchronize()
        assert out[1 - d].abs().max().item() == 0                        # the other direction's block is untouched
        assert rel_err(out[d].cpu(), ref) < 1e-4, (d, rel_err(out[d].cpu(), ref))


@pytest.mark.parametrize("M,V,chunk", [(300, 42, 4096), (5000, 10000, 2048), (130, 1000, 128), (64, 256, 4096)])
def test_linear_ce_head_fwd_bwd(pkg, monkeypatch, M, V, chunk):
    """Chunked Linear + cross-entropy(ignore_index=-1) + masked accuracy (ops.LinearCE) against fp64 torch: loss, accuracy,
    dX, dW, db; V = 42 exercises the pad-to-4 path, chunk < M the multi-chunk accumulation with a ragged last chunk."""
    monkeypatch.setattr(pkg.ops, "CE_CHUNK", chunk)
    rs = np.random.RandomState(M + V)
    x = torch.from_numpy(rs.standard_normal((M, 256)).astype(np.float32))
    w = torch.from_numpy((rs.standard_normal((V, 256)) / 16).astype(np.float32))
    b = torch.from_numpy(rs.standard_normal(V).astype(np.float32))
    y = torch.from_numpy(rs.randint(-1, V, size=M).astype(np.int64))
    y[:3] = -1
    xr, wr, br = (t.double().requires_grad_(True) for t in (x, w, b))
    logits = xr @ wr.t() + br
    ref = torch.nn.functional.cross_entropy(logits, y, ignore_index=-1)
    valid = y != -1
    ref_acc = (logits.max(1)[1][valid] == y[valid]).double().mean()
    (3.0 * ref).backward()
    xd, wd, bd = (t.cuda().requires_grad_(True) for t in (x, w, b))
    loss, acc = pkg.ops.linear_ce(xd, wd, bd, y.cuda())
    (3.0 * loss).backward()
    assert abs(loss.item() - ref.item()) < 2e-5 * abs(ref.item())
    assert abs(acc.item() - ref_acc.item()) < 1e-6
    assert rel_err(xd.grad.cpu(), xr.grad) < GRAD_TOL
    assert rel_err(wd.grad.cpu(), wr.grad) < GRAD_TOL
    assert rel_err(bd.grad.cpu(), br.grad) < GRAD_TOL
    with torch.no_grad():                                           # loss-only path (no gradient GEMMs)
        l2, a2 = pkg.ops.linear_ce(xd, wd, bd, y.cuda())
    assert abs(l2.item() - loss.item()) < 1e-6 * abs(loss.item()) and a2.item() == acc.item()
    lg = pkg.ops.LinearNT.apply(xd, wd, bd)                        # the posterior path (compute_posteriors), differentiable
    assert rel_err(lg.detach().cpu(), logits.detach()) < FWD_TOL
    xd.grad = None
    lg.square().sum().backward()
    xr.grad = None
    logits2 = xr @ wr.detach().t() + br.detach()
    logits2.square().sum().backward()
    assert rel_err(xd.grad.cpu(), xr.grad) < GRAD_TOL


def test_linear_ce_out_of_range_label_is_nan(pkg):
    x = torch.randn(8, 256, device="cuda"); w = torch.randn(44, 256, device="cuda"); b = torch.zeros(44, device="cuda")
    y = torch.tensor([0, 1, 2, 44, 3, -1, 5, 6], device="cuda")
    loss, _ = pkg.ops.linear_ce(x, w, b, y)
    assert torch.isnan(loss)


@pytest.mark.parametrize("impl", ["tc", "simt"])
@pytest.mark.parametrize("B,T,I,ds,p", [(5, 23, 256, 2, 0.5), (4, 50, 60, 1, 0.3), (17, 9, 256, 2, 0.5)])
def test_in_kernel_dropout_equals_the_explicit_canonical_mask(pkg, monkeypatch, impl, B, T, I, ds, p):
    """(p, seed) given to the GRU kernels = the mask slu_dropout_mask_gru writes for the same pair, given as a tensor: identical
    outputs and gradients, forward and backward (the backward kernel regenerates the mask instead of reading it)."""
    monkeypatch.setattr(pkg.ops, "GRU_IMPL", impl)
    rs = np.random.RandomState(B * 100 + T)
    gru = torch.nn.GRU(I, 128, batch_first=True, bidirectional=True).cuda()
    x = torch.from_numpy(rs.standard_normal((B, T, I)).astype(np.float32)).cuda()
    seed = 123456789012345
    mask = torch.empty(B, T, 256, device="cuda")
    pkg._lib.call("slu_dropout_mask_gru", mask.data_ptr(), B, T, p, seed, pkg._lib.stream())
    vals = torch.unique(mask).cpu().tolist()
    assert len(vals) == 2 and vals[0] == 0.0 and abs(vals[1] - 1.0 / (1.0 - p)) < 1e-6
    outs = []
    for drop in (mask, (p, seed)):
        xr = x.clone().requires_grad_(True)
        gru.zero_grad()
        y = pkg.ops.bigru(xr, gru, drop, ds)
        y.square().sum().backward()
        outs.append((y.detach().clone(), xr.grad.clone(), gru.weight_hh_l0.grad.clone(), gru.bias_ih_l0_reverse.grad.clone()))
    assert torch.equal(outs[0][0], outs[1][0])                       # same mask, same arithmetic: bit-identical forward
    for a, b in zip(outs[0][1:], outs[1][1:]):                       # gradients: split-K / atomic accumulation order varies
        assert rel_err(a, b) < 1e-5
    assert (outs[0][0] == 0).float().mean().item() > 0.02           # something was dropped


def test_canonical_dropout_mask_statistics(pkg):
    B, T, p = 8, 403, 0.5
    m = torch.empty(B, T, 256, device="cuda")
    pkg._lib.call("slu_dropout_mask_gru", m.data_ptr(), B, T, p, 42, pkg._lib.stream())
    keep = (m > 0).float()
    assert abs(keep.mean().item() - 0.5) < 4 * 0.5 / np.sqrt(m.numel())
    assert abs(keep.mean(dim=(0, 1)).sub(0.5).abs().max().item()) < 0.06          # per column
    assert abs(keep.mean(dim=(0, 2)).sub(0.5).abs().max().item()) < 0.06          # per time step
    k = keep.flatten()
    assert abs(((k[1:] - 0.5) * (k[:-1] - 0.5)).mean().item()) < 4 * 0.25 / np.sqrt(k.numel())    # neighbours uncorrelated
    m2 = torch.empty_like(m)
    pkg._lib.call("slu_dropout_mask_gru", m2.data_ptr(), B, T, p, 43, pkg._lib.stream())
    assert abs(((m2 > 0).float() * keep).mean().item() - 0.25) < 0.01              # different seeds: independent masks


@pytest.mark.parametrize("B,T", [(3, 12345), (300, 16000), (1, 81), (7, 64000)])
def test_sinc_persistent_kernel_equals_the_per_tile_kernel(pkg, B, T):
    """The persistent warp-specialised SincConv kernel (default) and the one-CTA-per-tile kernel issue the same MMAs in the same
    order on the same operand images: outputs and route bits must be identical, the reduced cut-off gradients equal to rounding
    (different summation order of the fp64 atomics); (300, 16000) gives every CTA several tiles and a ragged last wave."""
    p = R.synthetic_params()
    x = R.synthetic_batch(B, T, seed=T)[0].cuda()
    b1 = p[R.P + "phoneme_layers.0.filt_b1"].cuda(); band = p[R.P + "phoneme_layers.0.filt_band"].cuda()
    W = pkg.ops.sinc_filters(b1, band)
    L1 = (((T - 1) // 80 + 1) + 1) // 2
    gy = torch.randn(B, L1, 80, device="cuda")
    J = torch.empty(2, 80, 401, device="cuda")
    st = pkg._lib.stream()
    pkg._lib.call("slu_sinc_filters_jac", b1.data_ptr(), band.data_ptr(), J.data_ptr(), st)
    res = []
    try:
        for on in (1, 0):
            pkg._lib.load().slu_set_sinc_persistent(on)
            out = torch.empty(B, L1, 80, device="cuda"); route = torch.empty(B, L1, 80, device="cuda", dtype=torch.uint8)
            img = torch.empty(2 * 6 * 160 * 96, device="cuda", dtype=torch.bfloat16)
            d = torch.zeros(160, device="cuda", dtype=torch.float64)
            pkg._lib.call("slu_sincconv_fwd_tc", x.data_ptr(), W.data_ptr(), B, T, out.data_ptr(), route.data_ptr(), img.data_ptr(), st)
            pkg._lib.call("slu_sincconv_bwd_jac_tc", x.data_ptr(), gy.data_ptr(), route.data_ptr(), J.data_ptr(), B, T, d.data_ptr(),
                          img.data_ptr(), st)
            torch.cuda.synchronize()
            res.append((out, route, d))
    finally:
        pkg._lib.load().slu_set_sinc_persistent(1)
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    assert rel_err(res[0][2], res[1][2]) < 1e-6
