import hashlib
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "vq-vae-from-gaussian-vae_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(GOLDEN / f"{name}.npz", allow_pickle=False))


_CODEBOOKS = {}


def golden_codebook(g):
    """(codebook, nlp) fp32 arrays of a fixture; 65536-row codebooks are regenerated with the
    host generator and checked against the stored sha256."""
    if "codebook" in g:
        return g["codebook"], g["nlp"]
    key = (int(g["n_samples"]), int(g["group"]), int(g["seed"]))
    if key not in _CODEBOOKS:
        import torch  # noqa: F401
        from gq_b200.regularizer import prior_samples, _normal_log_prob
        cb = prior_samples(*key).float()
        nlp = _normal_log_prob(cb).float()
        _CODEBOOKS[key] = (cb.numpy(), nlp.numpy())
    cb, nlp = _CODEBOOKS[key]
    if hashlib.sha256(cb.tobytes()).hexdigest() != str(g["codebook_sha256"]):
        pytest.skip("prior_samples() on this host does not reproduce the fixture's codebook bits "
                    "(different torch/scipy build)")
    return cb, nlp


EVAL_CASES = [
    "eval_bchw_g16_n1024", "eval_blc_g16_n4096", "eval_bchw_g4_n4096", "eval_blc_g8_n4096",
    "eval_blc_g16_n4096_stress", "eval_bchw_g16_n2048_beta05", "eval_bchw_g16_n65536",
    "eval_bchw_g16_n65536_randinit", "eval_blc_g16_n3000_ragged",
]
TRAIN_CASES = ["train_blc_g16", "train_bchw_g4", "train_bchw_g16"]
