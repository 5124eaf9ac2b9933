"""Host-side behaviour of the drop-in module that needs no GPU: construction, buffers, the
lambda controller, error behaviour, the gq_cuda operator registration."""
import numpy as np
import pytest
import torch

from conftest import TRAIN_CASES, load_golden


def test_module_buffers_match_reference():
    import gq_b200
    g = load_golden("eval_bchw_g16_n1024")
    reg = gq_b200.GaussianQuantRegularizer(format="bchw", n_samples=1024, group=16, backend="cuda")
    assert np.array_equal(reg.prior_samples.numpy(), g["codebook"])        # gaussian.py:50
    assert np.array_equal(reg.normal_log_prob.numpy(), g["nlp"])           # gaussian.py:51-52
    assert reg.prior_samples.dtype == torch.float32
    # non-persistent buffers: nothing of the regularizer is checkpointed
    assert len(reg.state_dict()) == 0
    assert reg.log_n_samples == 10 and reg.lam == reg.lam_min == reg.lam_max == 1.0
    assert reg.training is True


def test_constructor_signature_matches_reference():
    import inspect
    import gq_b200
    sig = inspect.signature(gq_b200.GaussianQuantRegularizer.__init__)
    assert list(sig.parameters)[1:] == ["format", "n_samples", "group", "logvar_range", "tolerance", "lam_factor",
                                        "seed", "beta", "backend"]                          # gaussian.py:26-29
    d = {k: v.default for k, v in sig.parameters.items()}
    assert d["group"] == 1 and d["seed"] == 42 and d["beta"] == 1.0 and d["backend"] == "torch" and d["tolerance"] == 0.5
    sig2 = inspect.signature(gq_b200.GaussianQuantRegularizer2.__init__)
    assert list(sig2.parameters)[1:] == ["dim", "codebook_size", "dim_idx", "logvar_range", "tolerance", "lam_factor",
                                         "seed", "beta", "use_ste", "backend"]              # gaussian.py:182-184
    with pytest.raises(AssertionError):
        gq_b200.GaussianQuantRegularizer(format="nchw", n_samples=16)                      # gaussian.py:34


def test_no_cpu_path():
    import gq_b200
    reg = gq_b200.GaussianQuantRegularizer(format="blc", n_samples=256, group=16).eval()
    z = torch.zeros(1, 4, 32)
    with pytest.raises(NotImplementedError):
        reg(z)
    with pytest.raises(NotImplementedError):
        reg.train()(z)
    with pytest.raises(NotImplementedError):
        reg.dequant(torch.zeros(1, 4, 1, dtype=torch.int64))
    with pytest.raises(NotImplementedError):
        gq_b200.quantize_rows(None, torch.zeros(2, 16), torch.ones(2, 16))
    with pytest.raises(NotImplementedError):
        reg.eval().graphed_eval(z)                 # the graph replay has no CPU path either
    with pytest.raises(RuntimeError):
        reg.train().graphed_eval(z)                # and captures the eval branch only


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_lambda_controller_replays_reference(name):
    """gaussian.py:100-115 fed with the reference's own bits statistics reproduces its lambdas."""
    import gq_b200
    g = load_golden(name)
    reg = gq_b200.GaussianQuantRegularizer(format=str(g["format"]), n_samples=int(g["n_samples"]), group=int(g["group"]))
    for s in range(int(g["steps"])):
        np.testing.assert_allclose([reg.lam, reg.lam_min, reg.lam_max], g[f"lam_before{s}"], rtol=0)
        reg._update_lambdas(float(g[f"bits_mean{s}"]), float(g[f"bits_min{s}"]), float(g[f"bits_max{s}"]))
        np.testing.assert_allclose([reg.lam, reg.lam_min, reg.lam_max], g[f"lam_after{s}"], rtol=0)
        assert float(g[f"lam_info{s}"]) == pytest.approx(reg.lam, rel=1e-6)   # info["lam"] is post-update


def test_lambda_clamps():
    import gq_b200
    reg = gq_b200.GaussianQuantRegularizer(format="blc", n_samples=65536, group=16)
    for _ in range(2000):
        reg._update_lambdas(100.0, 0.0, 100.0)
    assert reg.lam_max == 1e3 and reg.lam_min == 1e-3
    for _ in range(3000):
        reg._update_lambdas(0.0, 16.0, 16.0)
    assert reg.lam_max == 1.0 and reg.lam_min == 1.0 and reg.lam == pytest.approx(1.01 ** -1000, rel=1e-9)


def test_gq_cuda_package_contract():
    """`import gq_cuda` must succeed (the reference's HAS_GQ_CUDA probe, gaussian.py:9-13) and expose
    ops.gq_cuda with the reference's signature, served by the op extension_cpp::gq."""
    import inspect
    import gq_cuda
    assert list(inspect.signature(gq_cuda.ops.gq_cuda).parameters) == ["a", "b", "c", "out", "d", "e", "f", "g"]
    schema = str(torch.ops.extension_cpp.gq.default._schema)
    assert "Tensor a, Tensor b, Tensor c, Tensor(a!) out, int d, int e, int f, float g" in schema
    with pytest.raises(NotImplementedError):        # no CPU kernel is registered, as in the reference
        gq_cuda.ops.gq_cuda(torch.zeros(2, 4), torch.ones(2, 4), torch.zeros(8, 4), torch.zeros(2, 8), 4, 2, 8, 1.0)


def test_configs_instantiate_unchanged():
    """The GQ configs' regularizer_config blocks (configs/sd3unet_gq_*.yaml, bsqvit_gq_*.yaml) resolve
    through the reference's own factory convention (pit/util.py:45-62: target + params) once
    pit.quantization.gaussian is bound to this package."""
    import importlib
    import sys
    import types
    import gq_b200
    pit = types.ModuleType("pit")
    q = types.ModuleType("pit.quantization")
    mod = types.ModuleType("pit.quantization.gaussian")
    mod.GaussianQuantRegularizer = gq_b200.GaussianQuantRegularizer
    mod.GaussianQuantRegularizer2 = gq_b200.GaussianQuantRegularizer2
    saved = {k: sys.modules.get(k) for k in ("pit", "pit.quantization", "pit.quantization.gaussian")}
    sys.modules.update({"pit": pit, "pit.quantization": q, "pit.quantization.gaussian": mod})
    try:
        blocks = [
            ("pit.quantization.gaussian.GaussianQuantRegularizer", dict(format="bchw", group=16, n_samples=4096, backend="cuda")),
            ("pit.quantization.gaussian.GaussianQuantRegularizer", dict(format="bchw", group=4, n_samples=4096, backend="cuda")),
            ("pit.quantization.gaussian.GaussianQuantRegularizer", dict(format="blc", group=8, n_samples=4096, backend="cuda")),
            ("pit.quantization.gaussian.GaussianQuantRegularizer", dict(format="bchw", group=16, n_samples=4096, seed=42, beta=1.0, backend="cuda")),
            ("pit.quantization.gaussian.GaussianQuantRegularizer2", dict(dim=16, codebook_size=4096, backend="cuda")),
        ]
        for target, params in blocks:
            module, cls = target.rsplit(".", 1)
            obj = getattr(importlib.import_module(module), cls)(**params)
            assert obj.prior_samples.shape[0] == 4096
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
