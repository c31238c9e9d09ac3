"""bring-up: size of the fp32-path deviation from the float64 oracle as the batch grows (external and compiled loss)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dib_oracle as O          # noqa: E402
from tests.test_gpu_parity import build_model, rel_err   # noqa: E402

for B in (256, 1024, 4160):
    rng = np.random.default_rng(11)
    cfg = O.DIBConfig([1] * 4, [128, 128], [256, 256], 8, feature_embedding_dimension=32, output_activation_fn="tanh")
    x = rng.standard_normal((B, 4)).astype(np.float32)
    y = rng.standard_normal((B, 8)).astype(np.float32)
    eps = rng.standard_normal((B, 4, 32)).astype(np.float32)
    for loss in ("external", "mse"):
        m = build_model(cfg, precision="fp32", loss=loss, seed=2)
        m.beta.assign(0.05)
        p = m.get_flat_weights()
        fr = O.forward(cfg, p, x, eps, 0.05)
        d_pred = (O.task_loss_grad("mse", fr.pred, y) / B).astype(np.float32)
        yy = d_pred if loss == "external" else y
        g, _ = m.compute_gradients(x, yy, eps=eps)
        g = g.cpu().numpy()
        g_ref, _ = O.train_grads(cfg, p, x, yy, eps, 0.05, loss)
        worst, off = [], 0
        for i, s in enumerate(cfg.param_shapes()):
            n = int(np.prod(s))
            worst.append((np.abs(g[off:off + n] - g_ref[off:off + n]).max() / np.abs(g_ref).max(), i, s))
            off += n
        worst.sort(reverse=True)
        print(f"B={B} loss={loss}: total {rel_err(g, g_ref):.2e}; worst vars", [(f"{w:.1e}", i, s) for w, i, s in worst[:3]], flush=True)
