"""Per-variable error of the tensor-core path against the float64 oracle on a committed golden case."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import dib_oracle as O
from tests.test_gpu_parity import LOSS_OF, build_model, load_case, make_labels, rel_err
name, mode = sys.argv[1], int(sys.argv[2])
cfg, z = load_case(os.path.join(os.path.dirname(__file__), "..", "tests", "golden"), name)
loss_name, loss = LOSS_OF[name]
m = build_model(cfg, precision="tf32", loss=loss_name)
m.debug_force_unfused(mode)
m.set_flat_weights(z["params"]); m.beta.assign(float(z["beta"]))
y = make_labels(np.random.default_rng(5), loss, z["x"].shape[0], cfg.output_dimensionality)
g, st = m.compute_gradients(z["x"], y, eps=z["eps"])
g_ref, fr = O.train_grads(cfg, z["params"], z["x"], y, z["eps"], float(z["beta"]), loss)
g = g.cpu().numpy()
names = []
for f in range(cfg.number_features):
    names += [f"f{f}.W0", f"f{f}.b0", f"f{f}.W1", f"f{f}.b1", f"f{f}.W2", f"f{f}.b2"]
names += [f"I.{k}{j}" for j in range(len(cfg.integration_network_architecture) + 1) for k in ("W", "b")]
off, rows = 0, []
for nm, s in zip(names, cfg.param_shapes()):
    n = int(np.prod(s)); rows.append((rel_err(g[off:off+n], g_ref[off:off+n]), nm)); off += n
print(f"{name} mode={mode} DIB_DBG={os.environ.get('DIB_DBG')} overall {rel_err(g, g_ref):.3e}  " +
      " ".join(f"{nm}:{e:.1e}" for e, nm in rows[-6:]) + "  enc worst " + str(max(rows[:-6])))
