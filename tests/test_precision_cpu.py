"""Host logic that needs no GPU: the precision-mode table (normalizing_flow/_engine.py), the host-side model of the split
bf16 pair layout the tests use (tests/split_pairs.py, include/nfdpm_b200.h NFDPM_BF16X2) and the loader of the unmodified
reference (oracle/reference_module.py, used by bench.py's reference arms and the drop-in test)."""
import os
import subprocess
import sys

import pytest
import torch

import split_pairs as SP
from normalizing_flow import _engine as E
from normalizing_flow import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_precision_mode_table(monkeypatch):
    want = {None: (N.SPLIT, torch.bfloat16), "auto": (N.SPLIT, torch.bfloat16), "fp32": (N.SPLIT, N.SPLIT),
            "bf16": (torch.bfloat16, torch.bfloat16), "fp32_simt": (torch.float32, torch.float32)}
    for mode, (infer, train) in want.items():
        if mode is None:
            monkeypatch.delenv("NFDPM_PRECISION", raising=False)
        else:
            monkeypatch.setenv("NFDPM_PRECISION", mode)
        assert E.coupling_dtype() == infer and E.coupling_dtype(train=True) == train, mode
    monkeypatch.setenv("NFDPM_PRECISION", "fp16")
    with pytest.raises(ValueError):
        E.precision()
    assert N._dt(torch.empty(1, dtype=N.SPLIT)) == N.BF16X2 == 3


def test_split_pair_host_model():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(37, 96, generator=g) * torch.logspace(-6, 3, 96)[None, :]
    enc = SP.encode(x)
    assert enc.dtype == N.SPLIT and enc.shape == x.shape                   # one 32-bit word per logical element
    hi, lo = SP.decode(enc, 37, 96)
    assert torch.equal(hi, x.bfloat16().float()) and torch.equal(lo, (x - hi).bfloat16().float())
    # layout: group g of 32 columns = 64 consecutive bf16, hi plane then lo plane
    raw = enc.view(torch.bfloat16).reshape(37, 192)
    assert torch.equal(raw[:, 64:96].float(), hi[:, 32:64]) and torch.equal(raw[:, 96:128].float(), lo[:, 32:64])
    err = (SP.value(enc, 37, 96) - x.double()).abs() / x.double().abs()
    assert float(err.max()) <= 2.0 ** -16                                   # vs 2^-8 for a single bf16


def test_reference_loader_imports_the_unmodified_package():
    sys.path.insert(0, ROOT)
    from oracle import reference_module as RM
    path = RM.find()
    if path is None:
        pytest.skip("no copy of the reference (baseline/_ref is staged by __graft_entry__.build())")
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from oracle import reference_module as RM\n"
            "nf = RM.import_reference()\n"
            "import torch\n"
            "f = nf.Glow(1, 2, 1)\n"
            "print(type(f).__module__, len(f.state_dict()), hasattr(nf, 'train'))\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 0, r.stderr[-2000:]
    mod, n_keys, has_train = r.stdout.split()[-3:]
    assert mod == "normalizing_flow.glow" and int(n_keys) > 20 and has_train == "True"
    # in THIS process the product's mirror of the same name is loaded: the loader must refuse instead of mixing them
    with pytest.raises(ImportError):
        RM.import_reference()
