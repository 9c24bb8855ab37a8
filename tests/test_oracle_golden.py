"""The oracle (oracle/glow_oracle.py) against the committed outputs of the unmodified
reference (tests/golden/*.npz, written by oracle/make_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import glow_oracle as O

CASES = ["glow_c1_L3_K2_b3_s32", "glow_c3_L3_K1_b2_s32", "glow_c3_L2_K1_b5_s16", "glow_c1_L2_K1_b2_s8_noprior"]


def _load(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    c, L, K, B, S, seed, lp = [int(v) for v in g["cfg"]]
    sd, psd = O.seeded_state(c, L, K, seed, learn_prior=bool(lp))
    # the regenerated weights are the ones the golden outputs were produced with
    assert np.allclose(np.array(O.state_checksum(sd)), g["checksum"], rtol=0, atol=1e-9)
    return g, (c, L, K, B, S, seed, bool(lp)), sd, psd


def test_all_fixtures_present(golden_dir):
    have = {os.path.basename(p)[:-4] for p in glob.glob(os.path.join(golden_dir, "*.npz"))}
    assert set(CASES) | {"glow_init_c1_L2_K1_b6_s16", "transforms", "glue"} <= have


@pytest.mark.parametrize("name", CASES)
def test_glow_forward_inverse(golden_dir, name):
    g, (c, L, K, B, S, seed, lp), sd, psd = _load(golden_dir, name)
    x = torch.from_numpy(g["x"])
    assert torch.equal(x, O.seeded_input((B, c, S, S), seed + 1000))
    ld = torch.zeros(B, dtype=torch.float64)
    logp = torch.zeros(B, dtype=torch.float64)
    zs, ld, logp = O.glow_transform(sd, x, L, K, ld, logp)
    assert len(zs) == L
    assert [tuple(z.shape[1:]) for z in zs] == O.output_shapes(L, c, S)
    for i, z in enumerate(zs):
        np.testing.assert_allclose(z.numpy(), g[f"z{i}"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ld.numpy(), g["ld"], rtol=1e-7)
    np.testing.assert_allclose(logp.numpy(), g["logp"], rtol=1e-7)
    ld2 = torch.zeros(B, dtype=torch.float64)
    _, ld2, none = O.glow_transform(sd, x, L, K, ld2, None)
    assert none is None
    np.testing.assert_allclose(ld2.numpy(), g["ld_nolp"], rtol=1e-7)
    gz = [torch.from_numpy(g[f"z{i}"]) for i in range(L)]
    np.testing.assert_allclose(O.glow_invert(sd, gz, L, K).numpy(), g["x_rec"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(O.glow_invert(sd, gz[-1:], L, K, temperature=0.0).numpy(), g["x_T0"],
                               rtol=1e-5, atol=2e-6)
    # north-star property: reconstruction error < 1e-4
    assert np.abs(g["x_rec"] - g["x"]).max() < 1e-4
    if lp:
        np.testing.assert_allclose(O.gaussian_prior_logp(psd, gz[-1]).numpy(), g["prior_logp"], rtol=1e-6)
        np.testing.assert_allclose(
            O.gaussian_prior_sample(psd, gz[-1].shape, 0.0, torch.zeros_like(gz[-1])).numpy(),
            g["prior_sample_T0"], rtol=1e-6, atol=1e-7)


def test_data_dependent_init(golden_dir):
    g = np.load(os.path.join(golden_dir, "glow_init_c1_L2_K1_b6_s16.npz"))
    c, L, K, B, S, seed, lp = [int(v) for v in g["cfg"]]
    sd, _ = O.seeded_state(c, L, K, seed, initialized=False)
    assert np.allclose(np.array(O.state_checksum(sd)), g["checksum"], atol=1e-9)
    x = torch.from_numpy(g["x"])
    ld = torch.zeros(B, dtype=torch.float64)
    logp = torch.zeros(B, dtype=torch.float64)
    zs, ld, logp = O.glow_transform(sd, x, L, K, ld, logp, init=True)
    for k in g.files:
        if k.startswith("sd/"):
            np.testing.assert_allclose(sd[k[3:]].numpy().reshape(g[k].shape), g[k], rtol=2e-5, atol=2e-6, err_msg=k)
    for i, z in enumerate(zs):
        np.testing.assert_allclose(z.numpy(), g[f"z{i}"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(ld.numpy(), g["ld"], rtol=1e-6)
    np.testing.assert_allclose(logp.numpy(), g["logp"], rtol=1e-6)


def test_transforms(golden_dir):
    g = np.load(os.path.join(golden_dir, "transforms.npz"))
    x = torch.from_numpy(g["x"])
    s, b = O.actnorm_stats(x)
    np.testing.assert_allclose(s.numpy(), g["scale"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(b.numpy(), g["bias"], rtol=1e-6, atol=1e-7)
    y, d = O.actnorm_fwd(x, s, b)
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose((torch.zeros(8) + d).numpy(), g["ld"], rtol=1e-6)
    np.testing.assert_allclose(O.actnorm_inv(y, s, b).numpy(), g["inv"], rtol=1e-6, atol=1e-6)
    # reference's own test properties (tests/transformations.py:29,35-41,63,85), EPS = 1e-3
    assert (O.actnorm_inv(y, s, b) - x).norm() < 1e-3
    assert y.mean(dim=(0, 2, 3)).abs().max() < 1e-3
    assert (y.var(dim=(0, 2, 3)) - 1).norm() < 1e-3
    w = torch.from_numpy(g["ic_w"])
    y, d = O.invconv_fwd(x, w)
    np.testing.assert_allclose(y.numpy(), g["ic_y"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose((torch.zeros(8) + d).numpy(), g["ic_ld"], rtol=1e-6)
    np.testing.assert_allclose(O.invconv_inv(y, w).numpy(), g["ic_inv"], rtol=1e-5, atol=1e-5)
    assert (O.invconv_inv(y, w) - x).norm() < 1e-3
    sdc, _ = O.seeded_state(1, 2, 1, 41)
    pre = "blocks.0.flows.0.affcoupling.net."
    xa = torch.from_numpy(g["ac_x"])
    y, d = O.coupling_fwd(xa, sdc, pre)
    np.testing.assert_allclose(y.numpy(), g["ac_y"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(d.numpy(), g["ac_ld"], rtol=1e-6)
    np.testing.assert_allclose(O.coupling_inv(y, sdc, pre).numpy(), g["ac_inv"], rtol=1e-5, atol=1e-5)
    assert (O.coupling_inv(y, sdc, pre) - xa).norm() < 1e-3
    xs = torch.from_numpy(g["sq_x"])
    assert np.array_equal(O.squeeze2x2(xs).numpy(), g["sq_y"])
    assert np.array_equal(O.unsqueeze2x2(torch.from_numpy(g["sq_y"])).numpy(), g["sq_inv"])
    assert np.array_equal(g["sq_inv"], g["sq_x"])
    # closed-form KAT: out channel = c*4 + h1*2 + w1
    assert O.squeeze2x2(xs)[1, 2 * 4 + 1 * 2 + 0, 1, 2] == xs[1, 2, 3, 4]


def test_glue(golden_dir):
    g = np.load(os.path.join(golden_dir, "glue.npz"))
    img = torch.from_numpy(g["img"])
    pre = O.preprocess_batch(img, 5, 32.0)
    assert np.array_equal(pre.numpy(), g["pre"])
    assert np.array_equal(O.postprocess_batch(pre, 32.0).numpy(), g["post"])
    np.testing.assert_allclose(O.bpd_loss(torch.from_numpy(g["ll"]), 32.0, 32 * 32 * 3.0).numpy(), g["loss"], rtol=1e-12)
    assert [tuple(r) for r in g["shapes"]] == O.output_shapes(3, 3, 32) == [(6, 16, 16), (12, 8, 8), (48, 4, 4)]


def test_closed_form_kats():
    """SURVEY §8c closed forms that need no oracle."""
    sd, psd = O.seeded_state(1, 2, 1, 5)
    pre = "blocks.0.flows.0."
    x = O.seeded_input((2, 4, 8, 8), 6)
    # zero-init coupling: scale = sigmoid(2), ld = (C/2)*P*log(sigmoid(2)+1e-6)
    for k in ("4.weight", "4.bias", "4.logs"):
        sd[pre + "affcoupling.net." + k] = torch.zeros_like(sd[pre + "affcoupling.net." + k])
    y, d = O.coupling_fwd(x, sd, pre + "affcoupling.net.")
    s2 = 1 / (1 + np.exp(-2.0))
    np.testing.assert_allclose(d.numpy(), 2 * 64 * np.log(s2 + 1e-6), rtol=1e-6)
    np.testing.assert_allclose(y[:, 2:].numpy(), x[:, 2:].numpy() * s2, rtol=1e-6)
    # gaussian logp at mean 0 / logs 0
    z = x[:, :2]
    np.testing.assert_allclose(O.gaussian_logp(z, torch.zeros_like(z), torch.zeros_like(z)).numpy(),
                               (-0.5 * z.numel() / 2 * np.log(2 * np.pi) - 0.5 * (z ** 2).reshape(2, -1).sum(1)).numpy(),
                               rtol=1e-6)


GRAD_CASES = ["glow_grad_c3_L2_K2_b3_s16", "glow_grad_c1_L3_K1_b2_s32"]


@pytest.mark.parametrize("name", GRAD_CASES)
def test_train_gradients(golden_dir, name):
    """Oracle loss and parameter gradients of one training step (trainer.py:154-164) against the unmodified
    reference's autograd: full tensors for small parameters, (norm, seeded projection) signatures for large ones."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    c, L, K, B, S, seed, _ = [int(v) for v in g["cfg"]]
    sd, psd = O.seeded_state(c, L, K, seed)
    assert np.allclose(np.array(O.state_checksum(sd)), g["checksum"], atol=1e-9)
    x = torch.from_numpy(g["x"])
    loss, gr, pgr = O.train_grads(sd, psd, x, L, K, 32.0, S * S * 3.0)
    assert abs(float(loss) - float(g["loss"])) < 1e-9
    for i, k in enumerate(g["names"]):
        k = str(k)
        t = pgr[k[len("prior/"):]] if k.startswith("prior/") else gr[k]
        nrm, proj = O.grad_signature(t, 5000 + i)
        ref_n, ref_p = g["sig/" + k]
        assert abs(nrm - ref_n) <= 1e-4 * ref_n + 1e-9, k
        assert abs(proj - ref_p) <= 1e-4 * max(ref_n, abs(ref_p)) * np.sqrt(t.numel()) ** 0 + 1e-6 * ref_n + 1e-9, k
        if ("grad/" + k) in g.files:
            assert np.allclose(t.numpy(), g["grad/" + k], rtol=1e-4, atol=1e-7), k


FORMAT_CASES = [("L3_c1_s32", 3, 1, 32, 2), ("L3_c3_s32", 3, 3, 32, 2), ("L4_c1_s32", 4, 1, 32, 1), ("L5_c3_s32", 5, 3, 32, 2),
                ("L5_c3_s64", 5, 3, 64, 1)]


def seeded_latents(dims, B, seed):
    rng = np.random.default_rng(seed)
    return [torch.from_numpy(rng.standard_normal((B,) + tuple(int(v) for v in d)).astype(np.float32)) for d in dims]


@pytest.mark.parametrize("idx", range(len(FORMAT_CASES)))
def test_cat_formater_oracle_against_reference(golden_dir, idx):
    """oracle.cat_format / cat_unformat against the unmodified reference CatFormater (tests/golden/formats.npz,
    oracle/make_golden_formats.py): bit-exact, both directions, L = 3, 4, 5."""
    g = np.load(os.path.join(golden_dir, "formats.npz"))
    name, L, c, S, B = FORMAT_CASES[idx]
    dims = g[name + "_dims"]
    assert [tuple(d) for d in dims] == O.output_shapes(L, c, S)
    lat = seeded_latents(dims, B, 700 + idx)
    cat = O.cat_format(lat)
    assert np.array_equal(cat.numpy(), g[name + "_cat"])
    for a, b in zip(O.cat_unformat(cat, dims), lat):
        assert torch.equal(a, b)
    rng = np.random.default_rng(800 + idx)
    q = torch.from_numpy(rng.standard_normal(tuple(cat.shape)).astype(np.float32))
    for j, t in enumerate(O.cat_unformat(q, dims)):
        assert np.array_equal(t.numpy(), g[f"{name}_post{j}"])


def test_pixel_formats_oracle_against_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "formats.npz"))
    xs = torch.from_numpy(g["post_x"])
    assert np.array_equal(O.postprocess_batch(xs, 32.0).numpy(), g["post_u8_32"])
    assert np.array_equal(O.postprocess_batch(xs, 256.0).numpy(), g["post_u8_256"])
    img, u = torch.from_numpy(g["pre_img"]), torch.from_numpy(g["pre_noise"])
    for n_bits in (5, 8, 3):
        assert np.array_equal(O.preprocess_batch(img, n_bits, 2.0 ** n_bits).numpy(), g[f"pre_{n_bits}"])
        assert np.array_equal(O.dequantize(img, n_bits, 2.0 ** n_bits, u).numpy(), g[f"dq_{n_bits}"])
