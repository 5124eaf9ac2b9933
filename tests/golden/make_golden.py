#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference module.

Run in the build container only (it imports /root/reference, which does not exist on the GPU
box):      python tests/golden/make_golden.py

The reference ships no golden vectors, so these fixtures -- outputs of
pit/quantization/gaussian.py executed on the CPU with fixed seeds -- are what pins the oracle
(oracle/gq_oracle.c) and, through it, the CUDA path.  Every array a test needs is stored, so
nothing has to be regenerated where the reference is absent; codebooks of 65536 rows are
stored as a sha256 plus the generator arguments (same image => same torch/scipy => same bits;
the test skips with a message if the hash differs).
"""
import hashlib
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
from pit.quantization.gaussian import (GaussianQuantRegularizer, GaussianQuantRegularizer2,  # noqa: E402
                                       prior_samples)

OUT = os.path.dirname(os.path.abspath(__file__))


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def latents(dist: str, shape_mu, gen: torch.Generator):
    """The three synthetic latent distributions of SURVEY.md 8(d)."""
    if dist == "trained":
        mu = torch.randn(shape_mu, generator=gen) * 0.866
        lv = -1.386 + 0.3 * torch.randn(shape_mu, generator=gen)
    elif dist == "randinit":
        mu = torch.randn(shape_mu, generator=gen) * 0.34
        lv = torch.randn(shape_mu, generator=gen) * 0.34
    elif dist == "stress":  # gq_cuda_extension/test/test_extension.py:18-19
        mu = torch.randn(shape_mu, generator=gen)
        lv = 2.0 * torch.log(torch.abs(torch.randn(shape_mu, generator=gen)) + 1e-3)
    else:
        raise ValueError(dist)
    return mu, lv


def rows_of(reg, z):
    """mu / std / log_std rows exactly as the reference builds them (gaussian.py:62-80,122-123)."""
    if reg.format == "bchw":
        b, c2, h, w = z.shape
        zz = z.permute(0, 2, 3, 1).reshape(b, h * w, c2)
    else:
        zz = z
    b, l, c2 = zz.shape
    c, g = c2 // 2, reg.group
    mu, lv = zz.chunk(2, 2)
    lv = torch.clamp(lv, reg.logvar_range[0], reg.logvar_range[1])
    std = torch.exp(0.5 * lv)
    mu_r = mu.reshape(b, l, g, c // g).permute(0, 1, 3, 2).reshape(-1, g).contiguous()
    std_r = std.reshape(b, l, g, c // g).permute(0, 1, 3, 2).reshape(-1, g).contiguous()
    return mu_r, std_r, std_r.log()


def top2_gap(reg, mu_r, std_r):
    from torch.distributions import Normal
    gaps = []
    for i in range(0, mu_r.shape[0], 64):
        q = Normal(mu_r[i:i + 64, None, :], std_r[i:i + 64, None, :])
        S = torch.sum(q.log_prob(reg.prior_samples[None]) - reg.normal_log_prob[None] * reg.beta, dim=2)
        t = torch.topk(S, 2, dim=1).values
        gaps.append(t[:, 0] - t[:, 1])
    return torch.cat(gaps)


def eval_case(name, fmt, b, spatial, C, g, N, dist, beta=1.0, seed=0, store_codebook=True):
    gen = torch.Generator().manual_seed(seed)
    reg = GaussianQuantRegularizer(format=fmt, n_samples=N, group=g, beta=beta).eval()
    if fmt == "bchw":
        h, w = spatial
        mu, lv = latents(dist, (b, C, h, w), gen)
        z = torch.cat([mu, lv], 1)
    else:
        mu, lv = latents(dist, (b, spatial, C), gen)
        z = torch.cat([mu, lv], 2)
    torch.manual_seed(1234)
    zhat, info = reg(z)
    dq = reg.dequant(info["indices"])
    mu_r, std_r, lstd_r = rows_of(reg, z)
    d = dict(
        format=fmt, group=g, n_samples=N, beta=beta, seed=42, z=z.numpy(), zhat=zhat.numpy(),
        indices=info["indices"].numpy(), zhat_noquant=info["zhat_noquant"].numpy(), dequant=dq.numpy(),
        mu_rows=mu_r.numpy(), std_rows=std_r.numpy(), log_std_rows=lstd_r.numpy(),
        gap=top2_gap(reg, mu_r, std_r).numpy(), rng_seed=1234,
        codebook_sha256=sha(reg.prior_samples.numpy()), nlp_sha256=sha(reg.normal_log_prob.numpy()),
    )
    if store_codebook:
        d["codebook"] = reg.prior_samples.numpy()
        d["nlp"] = reg.normal_log_prob.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "rows", mu_r.shape[0], "min gap", float(d["gap"].min()))


def train_case(name, fmt, b, spatial, C, g, N, dist, steps=3, seed=0):
    gen = torch.Generator().manual_seed(seed)
    reg = GaussianQuantRegularizer(format=fmt, n_samples=N, group=g).train()
    out = dict(format=fmt, group=g, n_samples=N, steps=steps)
    for s in range(steps):
        if fmt == "bchw":
            h, w = spatial
            mu, lv = latents(dist, (b, C, h, w), gen)
            z = torch.cat([mu, lv], 1)
        else:
            mu, lv = latents(dist, (b, spatial, C), gen)
            z = torch.cat([mu, lv], 2)
        if s == 1:  # exercise the logvar clamp and the three bands
            z = z.clone()
            z.view(-1)[::97] = 25.0
            z.view(-1)[1::89] = -35.0
        z.requires_grad_(True)
        torch.manual_seed(100 + s)
        lam0, lmin0, lmax0 = reg.lam, reg.lam_min, reg.lam_max
        zhat, info = reg(z)
        gen_w = torch.Generator().manual_seed(7 + s)
        wz = torch.randn(zhat.shape, generator=gen_w)
        loss = 0.1 * info["kl_loss"] + (zhat * wz).sum()  # discriminator_loss.py:321-325 weighting
        loss.backward()
        out.update({
            f"z{s}": z.detach().numpy(), f"zhat{s}": zhat.detach().numpy(), f"wz{s}": wz.numpy(),
            f"grad{s}": z.grad.numpy(), f"kl_loss{s}": info["kl_loss"].detach().numpy(),
            f"bits_mean{s}": info["bits-mean"].detach().numpy(), f"bits_min{s}": info["bits-min"].detach().numpy(),
            f"bits_max{s}": info["bits-max"].detach().numpy(), f"lam_info{s}": info["lam"].detach().numpy(),
            f"lam_before{s}": np.array([lam0, lmin0, lmax0]), f"lam_after{s}": np.array([reg.lam, reg.lam_min, reg.lam_max]),
            f"rng_seed{s}": 100 + s,
        })
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "kl_loss", [float(out[f"kl_loss{s}"]) for s in range(steps)])


def gq2_case(name, shape, dim, N, dim_idx, seed=0):
    gen = torch.Generator().manual_seed(seed)
    reg = GaussianQuantRegularizer2(dim=dim, codebook_size=N, dim_idx=dim_idx, backend="torch").eval()
    z = torch.randn(shape, generator=gen) * 0.6
    idx = [slice(None)] * len(shape)
    half = shape[dim_idx] // 2
    idx[dim_idx] = slice(half, None)
    z[tuple(idx)] = -1.386 + 0.3 * z[tuple(idx)]
    torch.manual_seed(55)
    zhat, info = reg(z)
    dq = reg.dequant(info["indices"])
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), dim=dim, n_samples=N, dim_idx=dim_idx, z=z.numpy(), zhat=zhat.detach().numpy(),
        indices=info["indices"].numpy(), zhat_quant=info["zhat_quant"].numpy(), zhat_noquant=info["zhat_noquant"].detach().numpy(),
        kl_loss=info["kl_loss"].detach().numpy(), bits_mean=info["bits-mean"].numpy(), bits_min=info["bits-min"].numpy(),
        bits_max=info["bits-max"].numpy(), lam_after=np.array([reg.lam, reg.lam_min, reg.lam_max]), dequant=dq.numpy(),
        codebook=reg.prior_samples.numpy(), nlp=reg.normal_log_prob.numpy(), rng_seed=55)
    print(name, "ok")


def legacy_case(name, M, N, dim, beta, seed=0):
    """The closed form the reference's own (uncalled) check defines: test_extension.py:5-6."""
    gen = torch.Generator().manual_seed(seed)
    noise = torch.randn(N, dim, generator=gen)
    mu = torch.randn(M, dim, generator=gen)
    std = torch.abs(torch.randn(M, dim, generator=gen)) + 0.05
    res = -torch.sum(((noise[None, :, :] - mu[:, None, :]) / std[:, None, :]) ** 2 - noise[None, :, :] ** 2 * beta, dim=2)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), mu=mu.numpy(), std=std.numpy(), noise=noise.numpy(),
                        beta=beta, result=res.numpy())
    print(name, "ok")


def usage_case(name, N, n, seed=0, batches=3):
    """Codebook usage / entropy exactly as eval.py computes them: all_hist = zeros(N) (eval.py:127);
    per batch `hist = torch.histogram(indices.cpu().float(), torch.range(0, N)); all_hist += hist[0]`
    (eval.py:152-154); then cal_ent (eval.py:137-141): usage = 1 - mean(hist == 0), p = hist / sum(hist),
    ent = -sum(p * log2(p + 1e-5)).  The index batches are synthetic (skewed, with unused codewords)."""
    import warnings
    gen = torch.Generator().manual_seed(seed)
    all_hist = torch.zeros([N])
    out = {"N": N}
    for bi in range(batches):
        # squared uniform -> heavy on low indices, and a block of codewords that is never used
        u = torch.rand(n, generator=gen)
        idx = (u * u * (N - N // 7)).long().clamp_(0, N - 1)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            hist = torch.histogram(idx.float(), torch.range(0, N))
        all_hist += hist[0]
        out[f"idx{bi}"] = idx.numpy()
    usage = torch.sum((all_hist == 0).to(dtype=torch.float32)) / all_hist.shape[0]
    p = all_hist / torch.sum(all_hist)
    ent = -torch.sum(p * torch.log2(p + 1e-5))
    out.update(batches=batches, hist=all_hist.numpy(), usage=float(1 - usage), entropy=float(ent))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "ok", float(1 - usage), float(ent))


if __name__ == "__main__":
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "usage":   # only the codebook-usage fixtures
        usage_case("usage_n65536", 65536, 60000)
        usage_case("usage_n1000", 1000, 5000, seed=1, batches=2)
        sys.exit(0)
    eval_case("eval_bchw_g16_n1024", "bchw", 1, (8, 8), 16, 16, 1024, "trained")
    eval_case("eval_blc_g16_n4096", "blc", 2, 32, 16, 16, 4096, "trained", seed=1)
    eval_case("eval_bchw_g4_n4096", "bchw", 2, (4, 4), 16, 4, 4096, "trained", seed=2)
    eval_case("eval_blc_g8_n4096", "blc", 1, 64, 16, 8, 4096, "randinit", seed=3)
    eval_case("eval_blc_g16_n4096_stress", "blc", 1, 64, 16, 16, 4096, "stress", seed=4)
    eval_case("eval_bchw_g16_n2048_beta05", "bchw", 1, (8, 8), 16, 16, 2048, "trained", beta=0.5, seed=5)
    eval_case("eval_bchw_g16_n65536", "bchw", 2, (8, 8), 16, 16, 65536, "trained", seed=6, store_codebook=False)
    eval_case("eval_bchw_g16_n65536_randinit", "bchw", 1, (8, 8), 16, 16, 65536, "randinit", seed=7, store_codebook=False)
    eval_case("eval_blc_g16_n3000_ragged", "blc", 1, 9, 16, 16, 3000, "trained", seed=8)
    train_case("train_blc_g16", "blc", 4, 64, 16, 16, 65536, "trained")
    train_case("train_bchw_g4", "bchw", 3, (8, 8), 16, 4, 65536, "trained", seed=1)
    train_case("train_bchw_g16", "bchw", 2, (8, 8), 16, 16, 65536, "randinit", seed=2)
    gq2_case("gq2_bchw_dim16", (2, 32, 4, 4), 16, 1024, 1)
    gq2_case("gq2_blc_dim8", (2, 16, 32), 8, 1024, 2, seed=1)
    legacy_case("legacy_b1", 24, 512, 16, 1.0)
    legacy_case("legacy_b07", 16, 256, 8, 0.7, seed=1)
    usage_case("usage_n65536", 65536, 60000)
    usage_case("usage_n1000", 1000, 5000, seed=1, batches=2)
