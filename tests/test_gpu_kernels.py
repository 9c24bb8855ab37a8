"""Kernel-level parity (GPU): every C-ABI entry point against a torch fp64/fp32 CPU restatement of the same
arithmetic on seeded inputs.  Tolerances are written per test; memory-layout kernels must be bit exact."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from normalizing_flow import _native as N
from oracle import glow_oracle as O
import split_pairs as SP

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0):
    g = np.random.default_rng(seed)
    return torch.from_numpy((scale * g.standard_normal(shape)).astype(np.float32))


def sync():
    torch.cuda.synchronize()


def a1f(t):
    """im2col rows of any operand format as fp32 values on the CPU."""
    if t.dtype == N.SPLIT:
        return SP.value(t.cpu(), t.shape[0], t.shape[1]).float()
    return t.float().cpu()


@pytest.mark.parametrize("C,P,B", [(3, 784, 2), (4, 256, 3), (8, 64, 5), (12, 256, 4), (16, 16, 9), (24, 64, 4),
                                   (48, 16, 7), (6, 49, 3), (96, 64, 2), (192, 16, 3), (12, 4096, 1)])
def test_channel_mix(C, P, B):
    x = rnd(B, C, P, seed=C + P)
    mt = rnd(C, C, seed=1)            # mt[i, o]
    beta = rnd(C, seed=2)
    y = torch.empty(B, C, P, device=DEV)
    N.channel_mix(x.cuda(), y, mt.cuda(), beta.cuda(), B, C, P, C * P, C * P)
    sync()
    ref = torch.einsum("io,bip->bop", mt.double(), x.double()) + beta.double()[None, :, None]
    assert torch.allclose(y.cpu().double(), ref, rtol=1e-5, atol=1e-5)


def test_channel_mix_strided_halves():
    # read the first half of a wider tensor, write into the second half of another (chunk / concat in place)
    B, C, P = 3, 8, 64
    wide = rnd(B, 2 * C, P, seed=3)
    out = torch.zeros(B, 2 * C, P, device=DEV)
    mt, beta = rnd(C, C, seed=4), rnd(C, seed=5)
    N.channel_mix(wide.cuda(), out.view(-1)[C * P:], mt.cuda(), beta.cuda(), B, C, P, 2 * C * P, 2 * C * P)
    sync()
    ref = torch.einsum("io,bip->bop", mt.double(), wide[:, :C].double()) + beta.double()[None, :, None]
    assert torch.allclose(out[:, C:].cpu().double(), ref, rtol=1e-5, atol=1e-5)
    assert out[:, :C].abs().sum() == 0


@pytest.mark.parametrize("C", [3, 4, 12, 24, 48, 96, 192])
def test_mix_prepare(C):
    g = np.random.default_rng(C)
    q, _ = np.linalg.qr(g.standard_normal((C, C)))
    w = torch.from_numpy((q + 0.1 * g.standard_normal((C, C))).astype(np.float32))
    s, b = rnd(C, seed=1, scale=0.3), rnd(C, seed=2, scale=0.3)
    f = dict(dtype=torch.float32, device=DEV)
    fwd_mt, fwd_beta, inv_mt, inv_beta, winv, logdet = (torch.empty(C * C, **f), torch.empty(C, **f),
                                                        torch.empty(C * C, **f), torch.empty(C, **f),
                                                        torch.empty(C * C, **f), torch.empty(1, **f))
    ws = torch.empty(2 * C * C, dtype=torch.float64, device=DEV)
    wd, sd_, bd = w.cuda(), s.cuda(), b.cuda()
    N.mix_prepare([N.MixItem(weight=wd.data_ptr(), scale=sd_.data_ptr(), bias=bd.data_ptr(), C=C, pad_=0,
                             fwd_mt=fwd_mt.data_ptr(), fwd_beta=fwd_beta.data_ptr(), inv_mt=inv_mt.data_ptr(),
                             inv_beta=inv_beta.data_ptr(), winv=winv.data_ptr(), logdet=logdet.data_ptr(),
                             lu_ws=ws.data_ptr())])
    sync()
    W = w.double()
    ref_ld = torch.slogdet(W)[1].float() + s.sum()
    assert abs(logdet.item() - ref_ld.item()) <= 1e-5 * max(1.0, abs(ref_ld.item()))
    Winv = torch.linalg.inv(W)
    assert torch.allclose(winv.cpu().reshape(C, C).double(), Winv, rtol=1e-4, atol=1e-5)
    ref_fwd = (W * torch.exp(s.double())[None, :]).T            # [i, o]
    assert torch.allclose(fwd_mt.cpu().reshape(C, C).double(), ref_fwd, rtol=1e-5, atol=1e-6)
    assert torch.allclose(fwd_beta.cpu().double(), ref_fwd.T @ b.double(), rtol=1e-4, atol=1e-5)
    ref_inv = (torch.exp(-s.double())[:, None] * Winv).T        # [i, o] = exp(-s[o]) Winv[o, i]
    assert torch.allclose(inv_mt.cpu().reshape(C, C).double(), ref_inv, rtol=1e-4, atol=1e-5)
    assert torch.equal(inv_beta.cpu(), -b)


def test_mix_prepare_identity_and_batch():
    # weight NULL -> identity; 40 items -> 3 launches of <=16
    C = 5
    f = dict(dtype=torch.float32, device=DEV)
    s = [rnd(C, seed=i, scale=0.2).cuda() for i in range(40)]
    outs = [torch.empty(1, **f) for _ in range(40)]
    mts = [torch.empty(C * C, **f) for _ in range(40)]
    N.mix_prepare([N.MixItem(weight=None, scale=s[i].data_ptr(), bias=None, C=C, pad_=0, fwd_mt=mts[i].data_ptr(),
                             fwd_beta=None, inv_mt=None, inv_beta=None, winv=None, logdet=outs[i].data_ptr(),
                             lu_ws=None) for i in range(40)])
    sync()
    for i in range(40):
        assert abs(outs[i].item() - s[i].sum().item()) < 1e-6
        assert torch.allclose(mts[i].cpu().reshape(C, C), torch.diag(torch.exp(s[i].cpu())), rtol=1e-6)


@pytest.mark.parametrize("B,C,H,W", [(2, 3, 4, 6), (3, 1, 32, 32), (2, 12, 16, 16), (1, 6, 8, 8)])
def test_squeeze_unsqueeze_exact(B, C, H, W):
    x = rnd(B, C, H, W, seed=7)
    y = torch.empty(B, 4 * C, H // 2, W // 2, device=DEV)
    N.squeeze(x.cuda(), y, B, C, H, W, C * H * W, C * H * W)
    back = torch.empty(B, C, H, W, device=DEV)
    N.unsqueeze(y, back, B, 4 * C, H // 2, W // 2, C * H * W, C * H * W)
    sync()
    assert torch.equal(y.cpu(), O.squeeze2x2(x))
    assert torch.equal(back.cpu(), x)


def test_squeeze_strided_and_copy_channels():
    B, C, H, W = 2, 4, 8, 8
    wide = rnd(B, 2 * C, H, W, seed=8)
    y = torch.empty(B, 4 * C, H // 2, W // 2, device=DEV)
    N.squeeze(wide.cuda(), y, B, C, H, W, 2 * C * H * W, C * H * W)        # squeeze the kept half in place
    z = torch.empty(B, C, H, W, device=DEV)
    N.copy_channels(wide.cuda().view(-1)[C * H * W:], z, B, C, H * W, 2 * C * H * W, C * H * W)
    sync()
    assert torch.equal(y.cpu(), O.squeeze2x2(wide[:, :C].contiguous()))
    assert torch.equal(z.cpu(), wide[:, C:])


@pytest.mark.parametrize("B,Cin,H,W,ld", [(2, 2, 16, 16, 64), (3, 6, 16, 16, 64), (2, 12, 8, 8, 128),
                                          (5, 24, 4, 4, 256), (1, 3, 7, 5, 32)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_im2col(B, Cin, H, W, ld, dt):
    x = rnd(B, 2 * Cin, H, W, seed=9)                     # use the first Cin channels of a wider tensor
    out = torch.full((B * H * W, ld), 7.0, dtype=dt, device=DEV)
    N.im2col3x3(x.cuda(), out, B, Cin, H, W, 2 * Cin * H * W, ld)
    sync()
    ref = F.unfold(x[:, :Cin], 3, padding=1).permute(0, 2, 1).reshape(B * H * W, Cin * 9)   # col = c*9 + ky*3 + kx
    got = out.cpu().float()
    assert torch.equal(got[:, :Cin * 9], ref.to(dt).float())
    assert got[:, Cin * 9:].abs().sum() == 0


def test_pack_matrix_layouts():
    C, Fe = 12, 64
    w3 = rnd(C, Fe, 3, 3, seed=10)
    ldp = 112
    out = torch.full((ldp, Fe), 5.0, device=DEV)
    N.pack_matrix(w3.cuda(), out, 9, C, Fe, 1, Fe * 9, 9, Fe, ldp)
    sync()
    ref = w3.permute(2, 3, 0, 1).reshape(9 * C, Fe)       # row = tap*C + co, col = ci
    assert torch.equal(out.cpu()[:9 * C], ref) and out.cpu()[9 * C:].abs().sum() == 0
    w1 = rnd(Fe, 6, 3, 3, seed=11)
    o1 = torch.full((Fe, 64), 5.0, dtype=torch.bfloat16, device=DEV)
    N.pack_matrix(w1.cuda(), o1, 1, Fe, 54, 0, 54, 1, 64, Fe)
    sync()
    assert torch.equal(o1.cpu().float()[:, :54], w1.reshape(Fe, 54).bfloat16().float())
    assert o1.cpu().float()[:, 54:].abs().sum() == 0


@pytest.mark.parametrize("M,N_,K", [(128, 128, 64), (300, 512, 64), (77, 112, 512), (1000, 12, 64), (129, 224, 128),
                                    (4096, 512, 512), (16, 432, 512), (257, 40, 16)])
@pytest.mark.parametrize("epi", [N.EPI_RAW, N.EPI_ACTNORM_RELU])
def test_gemm_nt_fp32(M, N_, K, epi):
    a, b = rnd(M, K, seed=M), rnd(N_, K, seed=N_ + 1)
    ldd = (N_ + 7) // 8 * 8
    d = torch.full((M, ldd), 3.0, device=DEV)
    es, eb = rnd(N_, seed=3, scale=0.2), rnd(N_, seed=4)
    N.gemm_nt(a.cuda(), K, b.cuda(), K, d, ldd, M, N_, K, epi, es.cuda(), eb.cuda())
    sync()
    ref = a.double() @ b.double().T
    if epi == N.EPI_ACTNORM_RELU:
        ref = torch.relu(torch.exp(es.double()) * (ref + eb.double()))
    got = d.cpu().double()
    assert torch.allclose(got[:, :N_], ref, rtol=1e-5, atol=1e-4 * math.sqrt(K) / 8)
    assert (got[:, N_:] == 3.0).all()            # padding columns untouched


def test_gemm_nt_fp32_bf16_out():
    M, N_, K = 200, 512, 64
    a, b = rnd(M, K, seed=1), rnd(N_, K, seed=2)
    d = torch.empty(M, N_, dtype=torch.bfloat16, device=DEV)
    N.gemm_nt(a.cuda(), K, b.cuda(), K, d, N_, M, N_, K)
    sync()
    ref = (a.double() @ b.double().T)
    assert torch.allclose(d.cpu().double(), ref, rtol=1e-2, atol=1e-2)


def _coupling_ref(pm_full, bias3, logs3, x, inverse):
    """pm_full: conv output [B, C, H, W] before bias; mirrors transforms.py:179-184 / 196-200."""
    C = x.shape[1]
    net = (pm_full + bias3.reshape(1, C, 1, 1)) * torch.exp(3 * logs3.reshape(1, C, 1, 1))
    log_s, t = net.chunk(2, 1)
    s = torch.sigmoid(log_s + 2.0)
    xa, xb = x.chunk(2, 1)
    if inverse:
        return torch.cat([xa, xb / (s + 1e-6) - t], 1), None
    return torch.cat([xa, (xb + t) * s], 1), torch.log(s + 1e-6).reshape(x.shape[0], -1).sum(1)


@pytest.mark.parametrize("B,C,H,W", [(3, 4, 28, 28), (5, 12, 16, 16), (6, 24, 8, 8), (19, 48, 4, 4), (2, 6, 5, 7),
                                     (2, 12, 64, 64)])
@pytest.mark.parametrize("inverse", [False, True])
def test_coupling_apply(B, C, H, W, inverse):
    Fe = 32
    h2 = rnd(B, Fe, H, W, seed=C)
    w3 = rnd(C, Fe, 3, 3, seed=C + 1, scale=0.05)
    bias3, logs3 = rnd(C, seed=2, scale=0.1), rnd(C, seed=3, scale=0.1)
    x = rnd(B, C, H, W, seed=4)
    # taps-as-N rows computed on the CPU: pm[m, tap*C+co] = sum_ci h2[m,ci] * w3[co,ci,tap]
    rows = h2.permute(0, 2, 3, 1).reshape(-1, Fe).double()
    wt = w3.permute(2, 3, 0, 1).reshape(9 * C, Fe).double()
    ldp = (9 * C + 15) // 16 * 16
    pm = torch.zeros(rows.shape[0], ldp)
    pm[:, :9 * C] = (rows @ wt.T).float()
    conv = F.conv2d(h2.double(), w3.double(), padding=1).float()
    ref_y, ref_ld = _coupling_ref(conv, bias3, logs3, x, inverse)
    P = H * W
    T = N.ld_tiles(P)
    part = torch.zeros(T * B, device=DEV)
    y = torch.empty(B, C, H, W, device=DEV)
    N.coupling_apply(pm.cuda(), ldp, bias3.cuda(), logs3.cuda(), x.cuda(), y, None if inverse else part, B, C, H, W,
                     C * P, C * P, inverse)
    sync()
    assert torch.allclose(y.cpu(), ref_y, rtol=2e-5, atol=2e-5)
    if not inverse:
        got = part.cpu().reshape(T, B).sum(0)
        assert torch.allclose(got, ref_ld, rtol=1e-5, atol=1e-4)
        # in place on the second half gives the same answer
        xin = x.cuda().clone()
        N.coupling_apply(pm.cuda(), ldp, bias3.cuda(), logs3.cuda(), xin, xin, part, B, C, H, W, C * P, C * P, False)
        sync()
        assert torch.equal(xin.cpu(), y.cpu())


@pytest.mark.parametrize("B,C,H,W", [(3, 4, 16, 16), (5, 12, 16, 16), (6, 24, 8, 8), (2, 8, 32, 32), (3, 6, 3, 5)])
def test_split_prior(B, C, H, W):
    P, Ch = H * W, C // 2
    x = rnd(B, C, H, W, seed=1)
    ldh = (C + 7) // 8 * 8
    h = torch.zeros(B * P, ldh)
    h[:, :C] = rnd(B * P, C, seed=2, scale=0.3)
    bias, logs = rnd(C, seed=3, scale=0.1), rnd(C, seed=4, scale=0.1)
    full = ((h[:, :C] + bias) * torch.exp(3 * logs)).reshape(B, H, W, C).permute(0, 3, 1, 2)
    mean, lg = full.chunk(2, 1)
    ref = O.gaussian_logp(x[:, Ch:], mean, lg)
    T = N.ld_tiles(P)
    part = torch.zeros(T * B, device=DEV)
    z = torch.empty(B, Ch, H, W, device=DEV)
    N.split_prior_logp(h.cuda(), ldh, bias.cuda(), logs.cuda(), x.cuda(), C * P, z, part, B, C, H, W)
    sync()
    assert torch.equal(z.cpu(), x[:, Ch:])
    assert torch.allclose(part.cpu().reshape(T, B).sum(0), ref, rtol=1e-5, atol=1e-3)
    # not-learned prior: standard normal
    N.split_prior_logp(None, 0, None, None, x.cuda(), C * P, None, part, B, C, H, W)
    sync()
    ref0 = O.gaussian_logp(x[:, Ch:], torch.zeros_like(mean), torch.zeros_like(lg))
    assert torch.allclose(part.cpu().reshape(T, B).sum(0), ref0, rtol=1e-5, atol=1e-3)
    # sampling writes mean + exp(lg)*T*eps into the second half only
    eps = rnd(B, Ch, H, W, seed=5)
    out = x.cuda().clone()
    N.split_prior_sample(h.cuda(), ldh, bias.cuda(), logs.cuda(), eps.cuda(), 0.7, out, C * P, B, C, H, W)
    sync()
    assert torch.equal(out.cpu()[:, :Ch], x[:, :Ch])
    assert torch.allclose(out.cpu()[:, Ch:], mean + torch.exp(lg) * 0.7 * eps, rtol=1e-5, atol=1e-6)


def test_gauss_const_and_accumulate():
    B, C, P = 7, 48, 16
    z = rnd(B, C, 4, 4, seed=1)
    bias, logs = rnd(2 * C, seed=2, scale=0.1), rnd(2 * C, seed=3, scale=0.1)
    psd = {"_GaussianPrior__conv.weight": torch.zeros(2 * C, 2 * C, 3, 3), "_GaussianPrior__conv.bias": bias,
           "_GaussianPrior__conv.logs": logs.reshape(1, -1, 1, 1)}
    ref = O.gaussian_prior_logp(psd, z)
    out = torch.empty(B, device=DEV)
    N.gauss_logp_const(z.cuda(), bias.cuda(), logs.cuda(), out, B, C, P)
    sync()
    assert torch.allclose(out.cpu(), ref, rtol=1e-5)
    eps = rnd(B, C, 4, 4, seed=4)
    smp = torch.empty(B, C, 4, 4, device=DEV)
    N.gauss_sample_const(eps.cuda(), bias.cuda(), logs.cuda(), 0.5, smp, B, C, P)
    sync()
    assert torch.allclose(smp.cpu(), O.gaussian_prior_sample(psd, z.shape, 0.5, eps), rtol=1e-5, atol=1e-6)
    # accumulate: fp64 and fp32 accumulators, partial rows + constants, in place
    part = rnd(3, B, seed=5)
    cval, cmul = rnd(4, seed=6), torch.tensor([256.0, 64.0, 16.0, 1.0])
    for dt in (torch.float64, torch.float32):
        acc = torch.arange(B, dtype=dt, device=DEV)
        ptr = acc.data_ptr()
        N.accumulate(acc, part.cuda(), 3, B, cval.cuda(), cmul.cuda(), 4)
        sync()
        ref_acc = torch.arange(B, dtype=torch.float64) + part.double().sum(0) + (cval.double() * cmul.double()).sum()
        assert acc.data_ptr() == ptr
        assert torch.allclose(acc.cpu().double(), ref_acc, rtol=1e-6 if dt == torch.float32 else 1e-12)


@pytest.mark.parametrize("layout", [0, 1])
def test_channel_stats(layout):
    B, C, H, W = 6, 5, 8, 8
    x = rnd(B, C, H, W, seed=1) * 3 + 10          # large mean/std ratio stresses the variance computation
    s_ref, b_ref = O.actnorm_stats(x)
    s, b = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    if layout == 0:
        N.channel_stats(x.cuda(), 0, B, C, H * W, C * H * W, s, b)
    else:
        rows = torch.zeros(B * H * W, 8)
        rows[:, :C] = x.permute(0, 2, 3, 1).reshape(-1, C)
        N.channel_stats(rows.cuda(), 1, B, C, H * W, 8, s, b)
    sync()
    assert torch.allclose(s.cpu(), s_ref.reshape(-1), rtol=1e-5, atol=1e-6)
    assert torch.allclose(b.cpu(), b_ref.reshape(-1), rtol=1e-6)


def test_layout_converters_and_actnorm_apply():
    B, C, H, W = 2, 5, 4, 6
    P = H * W
    x = rnd(B, C, H, W, seed=1)
    rows = torch.empty(B * P, 8, device=DEV)
    N.nchw_to_rows(x.cuda(), rows, B, C, P, C * P, 8)
    back = torch.empty(B, C, H, W, device=DEV)
    N.rows_to_nchw(rows, 8, 0, None, None, back, B, C, P)
    sync()
    assert torch.equal(back.cpu(), x)
    assert rows.cpu()[:, C:].abs().sum() == 0
    s, b = rnd(C, seed=2, scale=0.2), rnd(C, seed=3)
    N.rows_to_nchw(rows, 8, 2, s.cuda(), b.cuda(), back, B, C, P)
    y = torch.empty(B, C, H, W, device=DEV)
    N.actnorm_apply(x.cuda(), y, s.cuda(), b.cuda(), B, C, P, 0)
    inv = torch.empty(B, C, H, W, device=DEV)
    N.actnorm_apply(y, inv, s.cuda(), b.cuda(), B, C, P, 1)
    sync()
    ref = O.actnorm_fwd(x, s.reshape(C, 1, 1), b.reshape(C, 1, 1))[0]
    assert torch.allclose(y.cpu(), ref, rtol=1e-6, atol=1e-6) and torch.allclose(back.cpu(), ref, rtol=1e-6, atol=1e-6)
    assert torch.allclose(inv.cpu(), x, rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------ tcgen05 / TMEM / TMA bf16 GEMM
@pytest.mark.parametrize("M,N_,K", [(128, 256, 64), (300, 512, 64), (77, 112, 512), (4096, 224, 512), (2048, 432, 512),
                                    (129, 16, 128), (32768, 512, 512), (1000, 512, 256), (50000, 48, 64)])
@pytest.mark.parametrize("epi", [N.EPI_RAW, N.EPI_ACTNORM_RELU])
@pytest.mark.parametrize("out_dt", [torch.bfloat16, torch.float32])
def test_gemm_nt_bf16_tensor_core(M, N_, K, epi, out_dt):
    """bf16 operands, fp32 TMEM accumulation.  Reference = fp64 product of the SAME bf16-rounded operands, so the
    only error left is fp32 accumulation order (+ output rounding for bf16 stores: 2^-9 relative)."""
    a = rnd(M, K, seed=M % 97).bfloat16()
    b = (rnd(N_, K, seed=N_ + 1) * 0.1).bfloat16()
    d = torch.full((M, N_), 3.0, dtype=out_dt, device=DEV)
    es, eb = rnd(N_, seed=3, scale=0.2), rnd(N_, seed=4)
    N.gemm_nt(a.cuda(), K, b.cuda(), K, d, N_, M, N_, K, epi, es.cuda(), eb.cuda())
    sync()
    ref = a.double() @ b.double().T
    if epi == N.EPI_ACTNORM_RELU:
        ref = torch.relu(torch.exp(es.double()) * (ref + eb.double()))
    got = d.cpu().double()
    if out_dt == torch.float32:
        assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4)
    else:
        assert torch.allclose(got, ref, rtol=6e-3, atol=6e-3)


# ------------------------------------------------------------------ split bf16 pairs (NFDPM_BF16X2): fp32-faithful tensor-core mode
def test_split_pair_layout_pack_and_im2col():
    """pack_matrix / im2col3x3 writing split pairs == the host-side encoding of the same fp32 values, BIT EXACT
    (hi = bf16(v), lo = bf16(v - hi), groups of 32 columns interleaved), and hi + lo reproduces v to 2^-16."""
    C, Fe, ldp = 12, 64, 112
    w3 = rnd(C, Fe, 3, 3, seed=10)
    out = torch.full((ldp, Fe), 5, dtype=N.SPLIT, device=DEV)
    N.pack_matrix(w3.cuda(), out, 9, C, Fe, 1, Fe * 9, 9, Fe, ldp)
    ref = torch.zeros(ldp, Fe)
    ref[:9 * C] = w3.permute(2, 3, 0, 1).reshape(9 * C, Fe)
    sync()
    assert torch.equal(out.cpu(), SP.encode(ref))
    assert float((SP.value(out.cpu(), ldp, Fe) - ref.double()).abs().max()) <= 2.0 ** -16 * float(ref.abs().max())
    B, Cin, H, W, ld = 3, 6, 16, 16, 64
    x = rnd(B, 2 * Cin, H, W, seed=9)
    a = torch.full((B * H * W, ld), 7, dtype=N.SPLIT, device=DEV)
    N.im2col3x3(x.cuda(), a, B, Cin, H, W, 2 * Cin * H * W, ld)
    sync()
    cols = torch.zeros(B * H * W, ld)
    cols[:, :Cin * 9] = F.unfold(x[:, :Cin], 3, padding=1).permute(0, 2, 1).reshape(B * H * W, Cin * 9)
    assert torch.equal(a.cpu(), SP.encode(cols))


@pytest.mark.parametrize("M,N_,K", [(128, 256, 64), (300, 512, 64), (77, 112, 512), (4096, 224, 512), (2048, 432, 512),
                                    (129, 32, 128), (32768, 512, 512), (1000, 512, 256), (8192, 512, 128), (50000, 64, 64)])
@pytest.mark.parametrize("epi", [N.EPI_RAW, N.EPI_ACTNORM_RELU])
@pytest.mark.parametrize("out", ["f32", "split"])
def test_gemm_nt_split_pairs_reach_fp32_accuracy(M, N_, K, epi, out):
    """Split-pair operands, three tcgen05.mma per K slice (hi*hi + lo*hi + hi*lo), fp32 TMEM accumulation: the product of
    fp32 matrices within 2e-5 of the fp64 product relative to |A||B| row/column norms (bf16 operands: ~4e-3), i.e. the
    accuracy of an fp32 GEMM.  Split-pair OUTPUT carries the fp32 result to 2^-16 relative."""
    if out == "split" and N_ % 32:
        pytest.skip("split-pair rows hold whole groups of 32 columns")
    a, b = rnd(M, K, seed=M % 97), rnd(N_, K, seed=N_ + 1) * 0.1
    es, eb = rnd(N_, seed=3, scale=0.2), rnd(N_, seed=4)
    d = torch.full((M, N_), 3, dtype=torch.float32 if out == "f32" else N.SPLIT, device=DEV)
    N.gemm_nt(SP.encode(a).cuda(), K, SP.encode(b).cuda(), K, d, N_, M, N_, K, epi, es.cuda(), eb.cuda())
    sync()
    ref = a.double() @ b.double().T
    scale = a.double().norm(dim=1)[:, None] * b.double().norm(dim=1)[None, :]
    if epi == N.EPI_ACTNORM_RELU:
        ref = torch.relu(torch.exp(es.double()) * (ref + eb.double()))
        scale = torch.exp(es.double()) * (scale + eb.double().abs())
    got = d.cpu().double() if out == "f32" else SP.value(d.cpu(), M, N_)
    err = float(((got - ref).abs() / scale).max())
    assert err < 2e-5, err
    if out == "split":
        hi, lo = SP.decode(d.cpu(), M, N_)
        assert bool((lo.abs() <= 2.0 ** -8 * hi.abs()).all())      # lo is the rounding residual of hi


def test_gemm_nt_split_pairs_vs_plain_bf16_error():
    """The same product with plain bf16 operands is > 100x less accurate: what the mode buys (measured 4.8e-6 vs 2.3e-3
    relative L2; the split-pair figure is the 2^-18 rms representation error of hi + lo, twice)."""
    M, N_, K = 2048, 512, 512
    a, b = rnd(M, K, seed=1), rnd(N_, K, seed=2) * 0.1
    ref = a.double() @ b.double().T
    d3 = torch.empty(M, N_, device=DEV)
    N.gemm_nt(SP.encode(a).cuda(), K, SP.encode(b).cuda(), K, d3, N_, M, N_, K)
    d1 = torch.empty(M, N_, device=DEV)
    N.gemm_nt(a.bfloat16().cuda(), K, b.bfloat16().cuda(), K, d1, N_, M, N_, K)
    sync()
    e3 = float((d3.cpu().double() - ref).norm() / ref.norm())
    e1 = float((d1.cpu().double() - ref).norm() / ref.norm())
    assert e3 < 1e-5 and e1 > 100 * e3, (e3, e1)


# ------------------------------------------------------------------ fused step-boundary kernel
@pytest.mark.parametrize("B,C,H,W", [(3, 4, 16, 16), (5, 12, 16, 16), (4, 24, 8, 8), (7, 48, 4, 4), (2, 6, 5, 7), (2, 16, 2, 2)])
@pytest.mark.parametrize("a1_dt", [torch.float32, torch.bfloat16, N.SPLIT])
def test_flow_boundary_equals_unfused_chain(B, C, H, W, a1_dt):
    """nfdpm_flow_boundary == coupling_apply -> channel_mix -> im2col3x3 (forward and inverse), and
    squeeze -> channel_mix -> im2col3x3 at a level entry."""
    P, Ch = H * W, C // 2
    ldp = (9 * C + 15) // 16 * 16
    lda = (9 * Ch + 63) // 64 * 64
    x = rnd(B, C, H, W, seed=1).cuda()
    pm = (rnd(B * P, ldp, seed=2) * 0.1).cuda()
    bias3, logs3 = rnd(C, seed=3, scale=0.1).cuda(), rnd(C, seed=4, scale=0.1).cuda()
    mt, beta = rnd(C, C, seed=5, scale=0.4).cuda(), rnd(C, seed=6).cuda()
    for inverse in (False, True):
        # unfused chain
        y_ref = torch.empty_like(x)
        part_ref = torch.zeros(N.ld_tiles(P) * B, device=DEV)
        N.coupling_apply(pm, ldp, bias3, logs3, x, y_ref, None if inverse else part_ref, B, C, H, W, C * P, C * P, inverse)
        u_ref = torch.empty_like(x)
        N.channel_mix(y_ref, u_ref, mt, beta, B, C, P, C * P, C * P)
        a_ref = torch.full((B * P, lda), 9.0, dtype=a1_dt, device=DEV)
        N.im2col3x3(u_ref, a_ref, B, Ch, H, W, C * P, lda)
        # fused
        u = torch.empty_like(x)
        part = torch.zeros(B, device=DEV)
        a1 = torch.full((B * P, lda), 7.0, dtype=a1_dt, device=DEV)
        N.flow_boundary(x, C * P, False, pm, ldp, bias3, logs3, None if inverse else part, mt, beta, u, C * P, a1, lda,
                        B, C, H, W, inverse)
        sync()
        assert torch.allclose(u, u_ref, rtol=1e-5, atol=1e-5)
        assert torch.allclose(a1f(a1), a1f(a_ref), rtol=1e-2 if a1_dt == torch.bfloat16 else 1e-5, atol=1e-5)
        if not inverse:
            assert torch.allclose(part, part_ref.reshape(-1, B).sum(0), rtol=1e-5, atol=1e-4)
        # coupling only, in place, no mix / no im2col (last step of a level)
        xin = x.clone()
        N.flow_boundary(xin, C * P, False, pm, ldp, bias3, logs3, None, None, None, xin, C * P, None, 0, B, C, H, W, inverse)
        sync()
        assert torch.allclose(xin, y_ref, rtol=1e-6, atol=1e-6)
    # im2col only (entry of an inverse level)
    a_ref = torch.empty(B * P, lda, dtype=a1_dt, device=DEV)
    N.im2col3x3(x, a_ref, B, Ch, H, W, C * P, lda)
    a1 = torch.empty(B * P, lda, dtype=a1_dt, device=DEV)
    N.flow_boundary(x, C * P, False, None, 0, None, None, None, None, None, None, 0, a1, lda, B, C, H, W, False)
    sync()
    assert torch.equal(a1, a_ref)
    # level entry: squeeze the first C/4 channels of a wider, larger tensor, then mix + im2col
    if C % 4 == 0:
        src = rnd(B, C // 2, 2 * H, 2 * W, seed=7).cuda()                  # use its first C/4 channels
        sq = torch.empty(B, C, H, W, device=DEV)
        N.squeeze(src, sq, B, C // 4, 2 * H, 2 * W, (C // 2) * 4 * P, C * P)
        u_ref = torch.empty_like(sq)
        N.channel_mix(sq, u_ref, mt, beta, B, C, P, C * P, C * P)
        a_ref = torch.empty(B * P, lda, dtype=a1_dt, device=DEV)
        N.im2col3x3(u_ref, a_ref, B, Ch, H, W, C * P, lda)
        u = torch.empty_like(sq)
        a1 = torch.empty(B * P, lda, dtype=a1_dt, device=DEV)
        N.flow_boundary(src, (C // 2) * 4 * P, True, None, 0, None, None, None, mt, beta, u, C * P, a1, lda, B, C, H, W, False)
        sync()
        assert torch.allclose(u, u_ref, rtol=1e-5, atol=1e-5)
        assert torch.allclose(a1f(a1), a1f(a_ref), rtol=1e-2 if a1_dt == torch.bfloat16 else 1e-5, atol=1e-5)


def test_flow_boundary_rejects_large_images():
    assert N.flow_boundary_smem(24, 64, 64, 0, 0) > 200 * 1024
    x = torch.zeros(1, 24, 64, 64, device=DEV)
    with pytest.raises(RuntimeError, match="too large"):
        N.flow_boundary(x, 24 * 4096, False, None, 0, None, None, None, None, None, x, 24 * 4096, None, 0, 1, 24, 64, 64, False)


# ------------------------------------------------------------------ fused coupling network (tcgen05)
@pytest.mark.parametrize("M,N1,lda,N2,ldb", [(32768, 512, 512, 512, 512), (8192, 224, 256, 512, 512),
                                             (2048, 432, 448, 512, 512), (4096, 512, 512, 64, 64),
                                             (300, 112, 128, 512, 512), (64, 512, 512, 128, 128),
                                             (1000, 512, 512, 256, 256)])
def test_gemm_tn_bf16_tensor_core(M, N1, lda, N2, ldb):
    """D[N1,N2] = A[:, :N1]^T B[:, :N2] on tcgen05 with MN-major (transposed) operands read in place, split over M."""
    A = rnd(M, lda, seed=1).to(DEV).to(torch.bfloat16)
    B = rnd(M, ldb, seed=2).to(DEV).to(torch.bfloat16)
    D = torch.full((N1 * N2,), float("nan"), device=DEV)
    ws = torch.empty(N.gemm_tn_workspace(M, N1, N2), device=DEV)
    N.gemm_tn(A, lda, B, ldb, D, M, N1, N2, ws)
    sync()
    ref = A[:, :N1].double().T @ B[:, :N2].double()
    got = D.view(N1, N2).double()
    assert torch.isfinite(got).all()
    assert float((got - ref).norm() / ref.norm()) < 1e-5
    # accumulate=1 adds into D; the result is bitwise reproducible (fixed-order split reduction)
    D2 = D.clone()
    N.gemm_tn(A, lda, B, ldb, D2, M, N1, N2, ws, accumulate=True)
    D3 = D.clone()
    N.gemm_tn(A, lda, B, ldb, D3, M, N1, N2, ws, accumulate=True)
    sync()
    assert torch.equal(D2, D3)
    assert float((D2.view(N1, N2).double() - 2 * ref).norm() / ref.norm()) < 2e-5


@pytest.mark.parametrize("M,N1,lda,N2,ldb,mode,out_c", [(32768, 512, 512, 512, 512, 0, 0), (8192, 224, 256, 512, 512, 1, 24),
                                                       (2048, 432, 448, 512, 512, 1, 48), (4096, 512, 512, 128, 128, 2, 108),
                                                       (300, 112, 128, 512, 512, 1, 12), (1000, 512, 512, 256, 256, 0, 0),
                                                       (64, 512, 512, 64, 64, 2, 54)])
def test_gemm_tn_split_pairs(M, N1, lda, N2, ldb, mode, out_c):
    """Weight-gradient GEMM on split bf16 pairs: the accumulator holds hi*hi, hi*lo, lo*hi, lo*lo of every logical element
    and the in-kernel slab reduction adds them — the fp64 product of the fp32 operands to 1e-5 relative (plain bf16
    operands: ~3e-3), in the three output layouts (plain, ZeroConv weight [C,F,3,3], im2col padding stripped), bitwise
    reproducible."""
    A = torch.zeros(M, lda)
    A[:, :N1] = rnd(M, N1, seed=1)
    B = torch.zeros(M, ldb)
    B[:, :N2] = rnd(M, N2, seed=2)
    As, Bs = SP.encode(A).to(DEV), SP.encode(B).to(DEV)
    ref = A[:, :N1].double().T @ B[:, :N2].double()
    if mode == N.TN_OUT_TAPS:                       # n1 = tap*C + co -> [co][n2][tap]
        n_out = out_c * N2 * 9
        want = ref[:9 * out_c].reshape(9, out_c, N2).permute(1, 2, 0).reshape(-1)
    elif mode == N.TN_OUT_STRIP:
        n_out = N1 * out_c
        want = ref[:, :out_c].reshape(-1)
    else:
        n_out = N1 * N2
        want = ref.reshape(-1)
    D = torch.full((n_out,), float("nan"), device=DEV)
    ws = torch.empty(N.gemm_tn_workspace(M, N1, N2), device=DEV)
    N.gemm_tn(As, lda, Bs, ldb, D, M, N1, N2, ws, out_mode=mode, out_c=out_c)
    D2 = torch.full((n_out,), float("nan"), device=DEV)
    N.gemm_tn(As, lda, Bs, ldb, D2, M, N1, N2, ws, out_mode=mode, out_c=out_c)
    sync()
    got = D.double().cpu()
    assert torch.isfinite(got).all() and torch.equal(D, D2)
    assert float((got - want).norm() / want.norm()) < 1e-5


@pytest.mark.parametrize("B,C,H,W", [(3, 12, 16, 16), (5, 24, 8, 8), (4, 48, 4, 4), (2, 12, 32, 32)])
def test_coupling_bwd_operand_formats_agree(B, C, H, W):
    """nfdpm_coupling_bwd writes d(pm) as the K operand of the ZeroConv dgrad / wgrad GEMMs in the training format: the bf16
    and split-pair sinks must hold exactly the encodings of the fp32 sink's values (one-CTA-per-image and tiled forms), and
    every other output must not depend on the sink format."""
    P, Ch = H * W, C // 2
    ldp = (9 * C + 15) // 16 * 16
    Kp3 = (9 * C + 63) // 64 * 64
    M = B * P
    dy = rnd(B, C, P, seed=1).to(DEV)
    dld = rnd(B, seed=2).to(DEV)
    u = rnd(B, C, P, seed=3).to(DEV)
    pm = (rnd(M, ldp, seed=4) * 0.1).to(DEV)
    bias3, logs3 = rnd(C, seed=5, scale=0.1).to(DEV), rnd(C, seed=6, scale=0.1).to(DEV)
    T_c = N.coupling_bwd_tiles(C, H, W)
    outs = {}
    for dt in (torch.float32, torch.bfloat16, N.SPLIT):
        du = torch.zeros(B, C, P, device=DEV)
        dpm = torch.full((M, Kp3), 7, dtype=dt, device=DEV)
        dpar = torch.zeros(B * T_c * 2 * C, device=DEV)
        if T_c == 1:
            db, dl = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
            N.coupling_bwd(dy, C * P, dld, u, C * P, pm, ldp, bias3, logs3, du, C * P, dpm, Kp3, dpar, B, C, H, W, db, dl)
        else:
            N.coupling_bwd(dy, C * P, dld, u, C * P, pm, ldp, bias3, logs3, du, C * P, dpm, Kp3, dpar, B, C, H, W,
                           dp_scratch=torch.empty(M * C, device=DEV))
            db, dl = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
            N.reduce_rows2(dpar, db, dl, B * T_c, C, C, 2 * C)
        sync()
        outs[dt] = (du.cpu(), dpm.cpu(), db.cpu(), dl.cpu())
    du32, dpm32, db32, dl32 = outs[torch.float32]
    assert float(dpm32[:, 9 * C:].abs().max()) == 0.0 and float(dpm32.abs().max()) > 0
    for dt in (torch.bfloat16, N.SPLIT):
        du_, dpm_, db_, dl_ = outs[dt]
        assert torch.equal(du_, du32) and torch.equal(db_, db32) and torch.equal(dl_, dl32)
    assert torch.equal(outs[torch.bfloat16][1].float(), dpm32.bfloat16().float())
    assert torch.equal(outs[N.SPLIT][1], SP.encode(dpm32))


@pytest.mark.parametrize("M,Nn", [(300, 512), (4096, 512), (77, 64)])
def test_actnorm_relu_bwd_split_pairs(M, Nn):
    """(fp32 dh, split-pair h) -> split-pair dpre + column partials == the fp32 kernel on the decoded operands."""
    dh = rnd(M, Nn, seed=1).to(DEV)
    h = torch.relu(rnd(M, Nn, seed=2))
    scale = rnd(Nn, seed=3, scale=0.2).to(DEV)
    hs = SP.encode(h).to(DEV)
    h_val = SP.value(hs.cpu(), M, Nn).float().to(DEV)         # what the kernel sees: hi + lo
    rows = 32
    n_cta = (M + rows - 1) // rows
    dpre = torch.zeros(M, Nn, dtype=N.SPLIT, device=DEV)
    part = torch.zeros(n_cta * 2 * Nn, device=DEV)
    N.actnorm_relu_bwd(dh, Nn, hs, Nn, scale, dpre, Nn, part, M, Nn, rows)
    dpre_ref = torch.zeros(M, Nn, device=DEV)
    part_ref = torch.zeros(n_cta * 2 * Nn, device=DEV)
    N.actnorm_relu_bwd(dh, Nn, h_val, Nn, scale, dpre_ref, Nn, part_ref, M, Nn, rows)
    sync()
    assert torch.equal(dpre.cpu(), SP.encode(dpre_ref.cpu()))
    assert torch.allclose(part, part_ref, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("dta,dtb", [(torch.float32, torch.float32), (torch.bfloat16, torch.float32)])
def test_gemm_tn_cuda_core(dta, dtb):
    M, N1, N2 = 1500, 24, 112
    A = rnd(M, 32, seed=3).to(DEV).to(dta)
    B = rnd(M, 112, seed=4).to(DEV).to(dtb)
    D = torch.empty(N1 * N2, device=DEV)
    ws = torch.empty(N.gemm_tn_workspace(M, N1, N2), device=DEV)
    N.gemm_tn(A, 32, B, 112, D, M, N1, N2, ws)
    sync()
    ref = A[:, :N1].double().T @ B.double()
    assert float((D.view(N1, N2).double() - ref).norm() / ref.norm()) < 1e-5


# ------------------------------------------------------------------ ZeroConv GEMM + step boundary in one kernel
@pytest.mark.parametrize("B,C,H,W", [(5, 12, 16, 16), (3, 4, 16, 16), (5, 24, 8, 8), (4, 8, 8, 8), (11, 48, 4, 4),
                                     (8, 16, 4, 4), (6, 32, 4, 2)])
@pytest.mark.parametrize("a1_dt,op_dt", [(torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16),
                                         (N.SPLIT, N.SPLIT), (torch.float32, N.SPLIT)])
def test_gemm3_boundary_equals_gemm_plus_boundary(B, C, H, W, a1_dt, op_dt):
    """nfdpm_gemm3_boundary == nfdpm_gemm_nt (tcgen05, bf16 or split-pair operands, fp32 out) -> nfdpm_flow_boundary_stash:
    every sink (state, pre-mix stash, im2col rows, log-det partials, pm copy), forward and inverse, with and without the
    next mix; B deliberately not a multiple of the images-per-CTA count."""
    P, Ch, F = H * W, C // 2, 512
    ldp = (9 * C + 15) // 16 * 16
    lda = (9 * Ch + 63) // 64 * 64
    M = B * P
    assert N.gemm3_boundary_ok(B, C, H, W, F * (2 if op_dt == N.SPLIT else 1), ldp)
    x = rnd(B, C, H, W, seed=1).cuda()
    h2 = rnd(M, F, seed=2).clamp_min(0) * 0.5
    w3 = rnd(ldp, F, seed=3, scale=0.02)
    w3[9 * C:] = 0
    if op_dt == N.SPLIT:
        h2, w3 = SP.encode(h2).cuda(), SP.encode(w3).cuda()
    else:
        h2, w3 = h2.cuda().to(torch.bfloat16), w3.cuda().to(torch.bfloat16)
    bias3, logs3 = rnd(C, seed=4, scale=0.1).cuda(), rnd(C, seed=5, scale=0.1).cuda()
    mt, beta = rnd(C, C, seed=6, scale=0.4).cuda(), rnd(C, seed=7).cuda()
    pm_ref = torch.empty(M, ldp, device=DEV)
    N.gemm_nt(h2, F, w3, F, pm_ref, ldp, M, ldp, F)
    for inverse in (False, True):
        for with_mix in (True, False):
            m_, b_ = (mt, beta) if with_mix else (None, None)
            y_ref, xs_ref = torch.empty_like(x), torch.empty_like(x)
            a_ref = torch.full((M, lda), 3.0, dtype=a1_dt, device=DEV) if with_mix else None
            part_ref = torch.zeros(B, device=DEV)
            if inverse:
                N.flow_boundary(x, C * P, False, pm_ref, ldp, bias3, logs3, None, m_, b_, y_ref, C * P, a_ref,
                                lda if with_mix else 0, B, C, H, W, True)
            else:
                N.flow_boundary_stash(x, C * P, False, pm_ref, ldp, bias3, logs3, part_ref, m_, b_, y_ref, C * P, xs_ref,
                                      C * P, a_ref, lda if with_mix else 0, B, C, H, W)
            y, xs = torch.empty_like(x), torch.empty_like(x)
            a1 = torch.full((M, lda), 5.0, dtype=a1_dt, device=DEV) if with_mix else None
            part = torch.zeros(B, device=DEV)
            pm_out = torch.full((M, ldp), float("nan"), device=DEV)
            N.gemm3_boundary(h2, F, w3, pm_out, ldp, x, C * P, bias3, logs3, None if inverse else part, m_, b_, y, C * P,
                             None if inverse else xs, 0 if inverse else C * P, a1, lda if with_mix else 0, B, C, H, W, F,
                             ldp, inverse)
            sync()
            assert torch.equal(pm_out, pm_ref)
            assert torch.allclose(y, y_ref, rtol=1e-6, atol=1e-6), float((y - y_ref).abs().max())
            if with_mix:
                assert torch.allclose(a1f(a1), a1f(a_ref), rtol=1e-2 if a1_dt == torch.bfloat16 else 1e-6, atol=1e-6)
            if not inverse:
                assert torch.allclose(xs, xs_ref, rtol=1e-6, atol=1e-6)
                assert torch.allclose(part, part_ref, rtol=1e-6, atol=1e-5)
    # in place (y aliases the source), as Glow.transform runs it
    xin = x.clone()
    N.gemm3_boundary(h2, F, w3, None, 0, xin, C * P, bias3, logs3, None, mt, beta, xin, C * P, None, 0, None, 0, B, C, H, W,
                     F, ldp, False)
    y_ref = torch.empty_like(x)
    N.flow_boundary(x, C * P, False, pm_ref, ldp, bias3, logs3, None, mt, beta, y_ref, C * P, None, 0, B, C, H, W, False)
    sync()
    assert torch.allclose(xin, y_ref, rtol=1e-6, atol=1e-6)


def test_gemm3_boundary_support_query():
    assert not N.gemm3_boundary_ok(2, 12, 64, 64, 512, 112)      # image larger than a CTA tile
    assert not N.gemm3_boundary_ok(2, 6, 14, 14, 512, 64)        # 196 pixels: neither 256 nor a divisor of 128
    assert not N.gemm3_boundary_ok(2, 96, 8, 8, 512, 864)        # 9C > 512 TMEM columns
    assert not N.gemm3_boundary_ok(3, 16, 2, 2, 512, 144)        # 32 images per tile: more than one per warp
    assert N.gemm3_boundary_ok(128, 48, 4, 4, 512, 432)
