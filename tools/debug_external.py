"""bring-up: where does the tf32 external-loss step deviate from the oracle?  (per-variable errors over variants)"""
import itertools
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dib_oracle as O          # noqa: E402
from tests.test_gpu_parity import build_model, rel_err   # noqa: E402

rng = np.random.default_rng(11)
B = 256
x = rng.standard_normal((B, 4)).astype(np.float32)
eps = rng.standard_normal((B, 4, 32)).astype(np.float32)
for out, oact, beta, force in itertools.product((8, 6), (None, "tanh"), (0.05, 0.001), (0, 1, 3)):
    cfg = O.DIBConfig([1] * 4, [128, 128], [256, 256], out, feature_embedding_dimension=32, output_activation_fn=oact)
    d_pred = (np.random.default_rng(1).standard_normal((B, out)) / B).astype(np.float32)
    m = build_model(cfg, precision="tf32", loss="external", seed=2)
    if force:
        m.debug_force_unfused(force, batch_hint=B)
    m.beta.assign(beta)
    p = m.get_flat_weights()
    g, _ = m.compute_gradients(x, d_pred, eps=eps)
    g = g.cpu().numpy()
    g_ref, fr = O.train_grads(cfg, p, x, d_pred, eps, beta, "external")
    pred = np.asarray(m(x, eps=eps))
    worst, off = [], 0
    for i, s in enumerate(cfg.param_shapes()):
        n = int(np.prod(s))
        worst.append((np.abs(g[off:off + n] - g_ref[off:off + n]).max() / np.abs(g_ref).max(), i, s))
        off += n
    worst.sort(reverse=True)
    print(f"out={out} oact={oact} beta={beta} force={force}: total {rel_err(g, g_ref):.2e} pred {rel_err(pred, fr.pred):.2e} worst vars",
          [(f"{w:.1e}", i, s) for w, i, s in worst[:3]], flush=True)
