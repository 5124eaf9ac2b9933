"""Host-only helpers (no CUDA library is loaded by importing this module).

`prior_samples` is the reference's codebook generator kept verbatim in behaviour
(pit/quantization/gaussian.py:15-19): scrambled Sobol points through the normal inverse CDF,
float64 [n_samples, n_variable].  The codebook bits depend on the installed torch/scipy, so
parity is always asserted against a codebook produced in the same process."""
import torch
from scipy.stats import norm
from torch.distributions import Normal
from torch.quasirandom import SobolEngine


def prior_samples(n_samples: int, n_variable: int, seed_rec: int) -> torch.Tensor:
    sobol = SobolEngine(n_variable, scramble=True, seed=seed_rec)
    samples_sobol = sobol.draw(n_samples)
    return torch.from_numpy(norm.ppf(samples_sobol))


def normal_log_prob(x: torch.Tensor) -> torch.Tensor:
    """Normal(0, 1).log_prob(x) exactly as the reference constructor evaluates it
    (gaussian.py:51-52), on the host, once."""
    g = x.shape[-1]
    return Normal(torch.zeros([1, g]), torch.ones([1, g])).log_prob(x)


def host_codebook(n_samples: int, group: int, seed: int = 42):
    """(codebook, normal_log_prob) as fp32 numpy arrays, the module's two buffers (gaussian.py:50-52)."""
    cb = prior_samples(n_samples, group, seed).float()
    return cb.numpy(), normal_log_prob(cb).float().numpy()
