#!/usr/bin/env python
"""Where do the cycles of the int8 contraction go?  Runs one 16384-candidate chunk at N = 4096 with option "ozprof" and
prints, per launch mode, the clock64() sums of the MMA issuer / TMA producer / epilogue roles (gpk_get_oz_profile).
    python tools/oz_profile.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from robo_b200 import _lib                                   # noqa: E402
from robo_b200 import kernels as K                           # noqa: E402

N, D, M = 4096, 16, 16384
rng = np.random.RandomState(1234)
X = rng.rand(N, D)
y = np.sinc(X * 10 - 5).sum(axis=1) + 0.01 * rng.randn(N)
Xs = np.random.RandomState(4321).rand(M, D)
theta = np.concatenate(([0.0], np.full(D, np.log(D / 4.0))))
out = {}
for label, opts in (("default_cta_pair_two_passes", {}), ("cta_pair_two_passes_persistent", {"ozpersist": 1}),
                    ("persistent", {"ozpersist": 1, "ozpair": 0, "oztile": 64}), ("same_kernel_one_tile_per_cta", {"ozpersist": 2, "ozpair": 0, "oztile": 64}),
                    ("persistent_pair", {"ozpersist": 1, "ozpair": 1, "oztile": 64}), ("pair_one_tile_per_pair", {"ozpersist": 2, "ozpair": 1, "oztile": 64})):
    h = _lib.Handle(0)
    h.set_option("ozprof", 1)
    for k, v in opts.items():
        h.set_option(k, v)
    h.set_data(X, y)
    f = K.Product(K.ConstantKernel(theta[0], ndim=D), K.Matern52Kernel(np.exp(theta[1:]), ndim=D)).flatten()
    h.set_kernel(f["family"], f["log_amp"], f["axis"], f["group"], f["log_metric"])
    h.fit(1e-3 + 1.25e-12, float(np.mean(y)))
    for _ in range(3):
        r = h.acq(Xs, _lib.ACQ_EI, float(np.min(y)), 0.0, want_values=False)
    t = h.timings()
    p = h.oz_profile().astype(np.float64)
    lead = p[p[:, 6] > 0]                                    # CTAs that issued MMAs (all, or the pair leaders)
    tiles = lead[:, 6].sum()
    out[label] = {
        "vargemm_ms": t["vargemm_ms"], "ctas": int(p.shape[0]), "tiles": int(tiles), "best_idx": int(r["best_idx"]),
        "issuer_cycles_per_tile": lead[:, 0].sum() / tiles,
        "issuer_wait_operands_per_tile": lead[:, 1].sum() / tiles,
        "issuer_wait_tmem_drain_per_tile": lead[:, 2].sum() / tiles,
        "producer_wait_free_stage_per_tile": p[:, 3].sum() / max(1.0, tiles) / (p.shape[0] / lead.shape[0]),
        "epilogue_wait_accumulators_per_tile": p[:, 4].sum() / max(1.0, tiles) / (p.shape[0] / lead.shape[0]),
        "epilogue_drain_per_tile": p[:, 5].sum() / max(1.0, tiles) / (p.shape[0] / lead.shape[0]),
        "issuer_cycles_max_cta": float(lead[:, 0].max()), "issuer_cycles_min_cta": float(lead[:, 0].min()),
    }
    h.close()
print(json.dumps(out, indent=1))
