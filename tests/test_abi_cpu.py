"""CPU-side checks: the C-ABI library loads and exports every symbol include/nfdpm_b200.h declares, argument
validation fails loudly without touching a GPU, and the host-side mirror keeps the reference's state_dict layout."""
import ctypes
import os
import re

import pytest
import torch

import normalizing_flow as nf
from normalizing_flow import _native as N
from oracle import glow_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "nfdpm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nfdpm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(N.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nfdpm_b200.h but not exported"
    assert sorted(N.EXPORTS) == names, "python binding and header disagree"
    assert N.lib.nfdpm_version() == 100


def test_argument_validation_without_gpu():
    # bad arguments are rejected before any launch, with an explanatory message
    rc = N.lib.nfdpm_channel_mix(None, None, None, None, 1, 4, 16, 64, 64, None)
    assert rc != 0 and b"null pointer" in N.lib.nfdpm_last_error_string()
    rc = N.lib.nfdpm_squeeze(8, 8, 1, 1, 3, 4, 12, 12, None)
    assert rc != 0 and b"must be even" in N.lib.nfdpm_last_error_string()
    rc = N.lib.nfdpm_gemm_nt(8, 68, 8, 64, 8, 64, 4, 4, 64, 0, 0, 0, None, None, None)
    assert rc != 0 and b"multiples of 8" in N.lib.nfdpm_last_error_string()
    assert N.lib.nfdpm_ld_tiles(256) == 1 and N.lib.nfdpm_ld_tiles(257) == 2 and N.lib.nfdpm_ld_tiles(4096) == 16


def test_cpu_tensors_are_rejected_loudly():
    g = nf.Glow(1, 2, 1)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        g.transform(torch.zeros(2, 1, 8, 8), torch.zeros(2, dtype=torch.float64), None)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        nf.ActNorm(3).transform(torch.zeros(2, 3, 4, 4), torch.zeros(2), torch.zeros(2))


@pytest.mark.parametrize("c,L,K", [(1, 3, 2), (3, 3, 1), (3, 2, 2)])
def test_state_dict_layout_matches_reference(c, L, K):
    g = nf.Glow(c, L, K)
    want = O.glow_param_shapes(c, L, K)
    sd = g.state_dict()
    assert list(sd.keys()) == [k for k, _, _ in want]
    for k, shape, kind in want:
        assert tuple(sd[k].shape) == shape, k
        assert sd[k].dtype == (torch.uint8 if kind == "u8" else torch.float32), k
    seeded, psd = O.seeded_state(c, L, K, 3)
    g.load_state_dict(seeded, strict=True)
    gp = nf.GaussianPrior(2 ** (L + 1) * c)
    gp.load_state_dict(psd, strict=True)
    assert list(gp.state_dict().keys()) == ["_GaussianPrior__conv.weight", "_GaussianPrior__conv.bias",
                                            "_GaussianPrior__conv.logs"]
    nolp = nf.Glow(c, L, K, learn_prior_mean_logs=False)
    assert list(nolp.state_dict().keys()) == [k for k, _, _ in O.glow_param_shapes(c, L, K, learn_prior=False)]


def test_reference_init_semantics():
    torch.manual_seed(0)
    g = nf.Glow(3, 3, 2)
    sd = g.state_dict()
    w = sd["blocks.0.flows.0.invconv2d.weight"].reshape(12, 12)
    assert torch.allclose(w @ w.T, torch.eye(12), atol=1e-5)            # QR init (transforms.py:112-114)
    assert sd["blocks.0.flows.0.affcoupling.net.4.weight"].abs().sum() == 0   # ZeroConv2d (utils.py:37-38)
    assert sd["blocks.0.split.conv.bias"].abs().sum() == 0
    assert int(sd["blocks.0.flows.0.actnorm.is_initialized"]) == 0
    assert g.L == 3 and g.K == 2 and g.in_channel == 3 and hasattr(g, "device")
    assert len(g.blocks) == 2 and len(g.final_flows) == 2


def test_glue_matches_reference_vectors(golden_dir):
    import numpy as np
    g = np.load(os.path.join(golden_dir, "glue.npz"))
    img = torch.from_numpy(g["img"])
    pre = nf.preprocess_batch(img, 5, 32.0)
    assert np.array_equal(pre.numpy(), g["pre"])
    assert np.array_equal(nf.postprocess_batch(pre, 32.0).numpy(), g["post"])
    np.testing.assert_allclose(nf.calculate_loss(torch.from_numpy(g["ll"]), 32.0, 32 * 32 * 3.0).numpy(), g["loss"], rtol=1e-12)
    assert nf.calculate_output_shapes(3, 3, 32) == [(6, 16, 16), (12, 8, 8), (48, 4, 4)]
    with pytest.raises(ValueError):
        nf.calculate_output_shapes(3, 3, 30)
    a, b = nf.initialize_with_zeros(2, 5, torch.device("cpu"))
    assert a.dtype == torch.float64 and a.shape == (5,) and b is not a
    assert nf.get_item([1, 2], -3) is None and nf.get_item([1, 2], -2) == 1
