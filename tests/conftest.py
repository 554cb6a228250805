import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def sub(d, prefix):
    """{'a.b': v} -> {'b': tensor(v)} for keys under `prefix.`"""
    p = prefix + "."
    return {k[len(p):]: torch.from_numpy(np.array(v)) for k, v in d.items() if k.startswith(p)}


def rel_err(a, b):
    """max|a-b| / max|b|  -- the tolerance metric of this repo: north_star's "within 1e-5
    relative in fp32", normalised by the tensor's largest magnitude because summation order
    legitimately differs between implementations (element-wise relative error is undefined
    near zero)."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    den = b.abs().max().item()
    if den == 0.0:
        return (a - b).abs().max().item()
    return (a - b).abs().max().item() / den


# fp32 tolerance of the path (BASELINE.json north_star: logits and gradients within 1e-5 relative)
TOL = 1e-5


@pytest.fixture(params=["pair", "single"])
def umma_kernel(request):
    """Runs a tensor-core test twice: on the CTA-pair kernels (tcgen05 cta_group::2, the product path) and on the
    single-CTA kernels (kept for batches of one row tile and as the A/B reference)."""
    from gnn_tableextraction_b200 import _lib, ops

    v = 1 if request.param == "pair" else 0
    ops.set_tuning(_lib.GTE_TUNE_UMMA_PAIR, v)
    yield request.param
    ops.set_tuning(_lib.GTE_TUNE_UMMA_PAIR, 1)
