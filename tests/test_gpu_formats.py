"""Data formats either side of the flow (GPU): the CatFormater launch, device-side postprocess_batch and the
fused preprocess + dequantisation pass against the committed outputs of the unmodified reference (tests/golden/formats.npz)
and the CPU oracle.  Everything here is byte / permutation work: the bar is bit-exact."""
import os

import numpy as np
import pytest
import torch

import normalizing_flow as nf
from diffusion_prior import CatFormater, IdentityFormater, get_formater
from normalizing_flow import _native as N
from oracle import glow_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")
FORMAT_CASES = [("L3_c1_s32", 3, 1, 32, 2), ("L3_c3_s32", 3, 3, 32, 2), ("L4_c1_s32", 4, 1, 32, 1), ("L5_c3_s32", 5, 3, 32, 2),
                ("L5_c3_s64", 5, 3, 64, 1)]


def seeded_latents(dims, B, seed):
    rng = np.random.default_rng(seed)
    return [torch.from_numpy(rng.standard_normal((B,) + tuple(int(v) for v in d)).astype(np.float32)) for d in dims]


@pytest.mark.parametrize("idx", range(len(FORMAT_CASES)))
def test_cat_formater_against_reference_golden(golden_dir, idx):
    g = np.load(os.path.join(golden_dir, "formats.npz"))
    name, L, c, S, B = FORMAT_CASES[idx]
    fm = get_formater("CatFormater")(L, c, S)
    assert isinstance(fm, CatFormater) and fm.get_num_latent_parts() == 1
    assert np.array_equal(np.array(fm.latent_dims), g[name + "_dims"])
    assert np.array_equal(np.array(fm.get_input_shapes()), g[name + "_input_shapes"])
    lat = [t.to(DEV) for t in seeded_latents(fm.latent_dims, B, 700 + idx)]
    l0 = N.launch_count
    out = fm.process_latents(lat)
    assert N.launch_count - l0 == 1 and len(out) == 1          # ONE launch for all parts
    assert np.array_equal(out[0].cpu().numpy(), g[name + "_cat"])
    back = fm.postprocess(out)
    assert len(back) == L and all(torch.equal(a, b) for a, b in zip(back, lat))
    rng = np.random.default_rng(800 + idx)
    q = torch.from_numpy(rng.standard_normal(tuple(out[0].shape)).astype(np.float32)).to(DEV)
    for j, t in enumerate(fm.postprocess([q])):
        assert np.array_equal(t.cpu().numpy(), g[f"{name}_post{j}"])
    idf = get_formater("IdentityFormater")(L, c, S)
    assert isinstance(idf, IdentityFormater) and idf.get_num_latent_parts() == L
    assert np.array_equal(np.array(idf.get_input_shapes()), g[name + "_id_shapes"])
    assert idf.process_latents(lat) is lat and idf.postprocess(lat) is lat
    with pytest.raises(ValueError, match="Invalid formater name"):      # reference latent_formaters.py:261-262
        get_formater("nope")


def test_cat_formater_full_size_and_oracle(monkeypatch):
    monkeypatch.setenv("NFDPM_PRECISION", "fp32")
    """BASELINE config 5 size (1024 latents of the L3 MNIST flow) and config 4 geometry (L5, 128x128): equals the oracle,
    round trip is the identity, and the decode chain postprocess -> Glow.sample -> postprocess_batch runs end to end."""
    for L, c, S, B in ((3, 1, 32, 1024), (5, 3, 128, 8)):
        fm = CatFormater(L, c, S)
        lat = seeded_latents(fm.latent_dims, B, 5)
        cat = fm.process_latents([t.to(DEV) for t in lat])[0]
        assert torch.equal(cat.cpu(), O.cat_format(lat))
        for a, b in zip(fm.postprocess([cat]), lat):
            assert torch.equal(a.cpu(), b)
    sd, _ = O.seeded_state(1, 3, 2, 9)
    flow = nf.Glow(1, 3, 2).to(DEV)
    flow.load_state_dict(sd)
    fm = CatFormater(3, 1, 32)
    rng = np.random.default_rng(11)
    q = torch.from_numpy(rng.standard_normal((16, 16, 8, 8)).astype(np.float32))
    img = flow.sample(fm.postprocess([q.to(DEV)]), postprocess_func=lambda t: nf.postprocess_batch(t, 32.0))
    assert img.dtype == torch.uint8 and img.device.type == "cpu" and img.shape == (16, 1, 32, 32)
    with torch.no_grad():
        xo = O.glow_invert(sd, O.cat_unformat(q, fm.latent_dims), 3, 2)
    want = O.postprocess_batch(xo, 32.0)
    # the flow itself is fp32 arithmetic in another summation order: a value within 1e-5 of a bin edge may land in the
    # neighbouring bin; everything else must be identical
    diff = (img.int() - want.int()).abs()
    assert int((diff > 0).sum()) <= 0.001 * diff.numel() and int(diff.max()) <= 8


def test_cat_formater_autograd_and_errors():
    fm = CatFormater(3, 3, 32)
    lat = [t.to(DEV).requires_grad_(True) for t in seeded_latents(fm.latent_dims, 2, 3)]
    cat = fm.process_latents(lat)[0]
    w = torch.from_numpy(np.random.default_rng(4).standard_normal(tuple(cat.shape)).astype(np.float32)).to(DEV)
    (cat * w).sum().backward()
    for t, gw in zip(lat, O.cat_unformat(w.cpu(), fm.latent_dims)):       # d/dlatent = the inverse permutation of w
        assert torch.equal(t.grad.cpu(), gw)
    q = w.clone().requires_grad_(True)
    parts = fm.postprocess([q])
    ws = [torch.full_like(p, float(i + 1)) for i, p in enumerate(parts)]
    sum((p * v).sum() for p, v in zip(parts, ws)).backward()
    assert torch.equal(q.grad.cpu(), O.cat_format([v.cpu() for v in ws]))
    with pytest.raises(ValueError):
        fm.process_latents(lat[:2])
    with pytest.raises(ValueError):
        fm.process_latents([lat[0], lat[1], lat[2][:, :4]])
    with pytest.raises(AssertionError):
        fm.postprocess([cat, cat])
    with pytest.raises(ValueError):
        fm.postprocess([cat[:, :5]])
    with pytest.raises((RuntimeError, TypeError, ValueError)):
        fm.process_latents([t.detach().cpu() for t in lat])                 # no host path
    with pytest.raises(RuntimeError):                                        # the C ABI validates the part table itself
        N.latent_format([(lat[0].detach(), 1, 0, 5)], cat.detach(), 2, 48, 8, 8, True)


def test_pixel_formats_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "formats.npz"))
    xs = torch.from_numpy(g["post_x"]).to(DEV)
    for n_bins, key in ((32.0, "post_u8_32"), (256.0, "post_u8_256")):
        out = nf.postprocess_batch(xs, n_bins)
        assert out.dtype == torch.uint8 and out.device.type == "cpu"
        assert np.array_equal(out.numpy(), g[key])
    img, u = torch.from_numpy(g["pre_img"]).to(DEV), torch.from_numpy(g["pre_noise"]).to(DEV)
    for n_bits in (5, 8, 3):
        l0 = N.launch_count
        pre = nf.preprocess_batch(img, n_bits, 2.0 ** n_bits)
        dq = nf.preprocess_batch(img, n_bits, 2.0 ** n_bits, noise=u)
        assert N.launch_count - l0 == 2
        assert np.array_equal(pre.cpu().numpy(), g[f"pre_{n_bits}"])
        assert np.array_equal(dq.cpu().numpy(), g[f"dq_{n_bits}"])
    # full size (config 2 batch, beyond one wave) against the oracle, odd length tail, integer n_bins like the trainers pass
    rng = np.random.default_rng(6)
    big = torch.from_numpy(rng.random((128, 3, 32, 32)).astype(np.float32))
    nz = torch.from_numpy(rng.random((128, 3, 32, 32)).astype(np.float32))
    assert torch.equal(nf.preprocess_batch(big.to(DEV), 5, 32, noise=nz.to(DEV)).cpu(), O.dequantize(big, 5, 32.0, nz))
    m = torch.from_numpy((rng.random(100003) * 1.6 - 0.8).astype(np.float32))
    assert torch.equal(nf.postprocess_batch(m.to(DEV), 32), O.postprocess_batch(m, 32.0))
    bad = torch.tensor([float("nan"), float("inf"), -float("inf"), 0.25] * 4, device=DEV)
    assert nf.postprocess_batch(bad, 32.0).tolist() == [0, 255, 0, 192] * 4
