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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def load_golden(name):
    """Returns dict with meta + torch tensors grouped by prefix."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {"param": {}, "gparam": {}, "in": {}, "gin": {}, "out": {}, "probe": {}, "mask": {}}
    for k in z.files:
        if "/" in k:
            grp, nm = k.split("/", 1)
            out[grp][nm] = torch.from_numpy(z[k].copy())
        else:
            out[k] = z[k]
    C, n_head, L, A, S, B, scale = [int(v) for v in out["meta"]]
    out["cfg"] = dict(C=C, n_head=n_head, L=L, A=A, S=S, B=B, scale=scale)
    return out


GPT_CASES = ["gpt_tiny", "gpt_c64_t962"]
STAGE_CASES = ["stage_tiny_s1", "stage_tiny_s2", "stage_tiny_s4", "stage_tiny_s8"]


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def assert_close(a, b, rtol, atol=1e-7, msg=""):
    """||a-b|| <= rtol*||b|| + atol*sqrt(numel).  The atol term covers gradients that are
    mathematically zero (e.g. attn.key.bias: softmax is invariant to a key-bias shift)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, "%s shape %s vs %s" % (msg, tuple(a.shape), tuple(b.shape))
    err = float((a - b).norm())
    bound = rtol * float(b.norm()) + atol * (b.numel() ** 0.5)
    assert err <= bound, "%s: err %.3e > bound %.3e (rel %.3e)" % (msg, err, bound, err / (float(b.norm()) + 1e-30))


@pytest.fixture(scope="session")
def cuda_dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
