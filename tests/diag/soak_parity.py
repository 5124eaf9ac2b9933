#!/usr/bin/env python
"""Parity soak: every row of large batches against the CPU oracle (bit-exact indices and codewords), for the
groups and latent distributions of SURVEY 8(d) and both search modes.  The north-star bar is a near-tie
mismatch rate < 1e-6 and zero mismatches on non-ties; this run checks ~1.5 M rows.  Prints one JSON object.

    python tests/diag/soak_parity.py [rows_per_case]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "vq-vae-from-gaussian-vae_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "diag"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import gq_b200  # noqa: E402
import oracle  # noqa: E402
from gq_b200_host import host_codebook  # noqa: E402
from sweep import rows  # noqa: E402


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    oracle.set_num_threads(0)
    dev = torch.device("cuda:0")
    cases = [(16, "trained", "auto"), (16, "randinit", "auto"), (16, "stress", "fast"), (16, "stress", "robust"),
             (16, "trained", "robust"), (8, "trained", "auto"), (8, "stress", "robust"), (4, "trained", "auto"),
             (4, "stress", "auto"), (32, "trained", "auto"), (2, "trained", "auto")]
    out = {"rows_per_case": M, "N": 65536, "cases": []}
    total = bad_total = 0
    for g, dist, mode in cases:
        m_case = M if g >= 8 else M // 2
        if dist == "stress" and mode == "fast":
            m_case = min(m_case, 16384)          # the exact-scan fallback handles most of these rows
        cbn, nlp = host_codebook(65536, g)
        cb = gq_b200.Codebook(torch.from_numpy(cbn).to(dev), torch.from_numpy(nlp).to(dev), 1.0)
        cb.set_mode(mode)
        mu, std, lstd = rows(m_case, g, 1000 + g, dist)
        stats = torch.zeros(8, dtype=torch.int32, device=dev)
        idx, zhat = gq_b200.quantize_rows(cb, mu.to(dev), std.to(dev), lstd.to(dev), stats=stats)
        t0 = time.time()
        o_idx, o_zhat, _, gap = oracle.quantize(mu.numpy(), std.numpy(), lstd.numpy(), cbn, nlp)
        got = idx.cpu().numpy()
        bad = np.nonzero(got != o_idx)[0]
        zeq = bool(np.array_equal(zhat.cpu().numpy()[got == o_idx], o_zhat[got == o_idx]))
        rec = {"g": g, "latents": dist, "mode": mode, "rows": m_case, "index_mismatches": int(len(bad)),
               "mismatch_top2_gaps": [float(x) for x in gap[bad][:8]], "codewords_equal": zeq,
               "rows_to_exact_scan": int(stats[2]), "rows_settled": int(stats[4]), "oracle_s": round(time.time() - t0, 1)}
        out["cases"].append(rec)
        print(rec, file=sys.stderr, flush=True)
        total += m_case
        bad_total += len(bad)
    out["rows_total"] = total
    out["index_mismatches_total"] = bad_total
    print(json.dumps(out))


if __name__ == "__main__":
    main()
