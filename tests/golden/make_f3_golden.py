"""Golden vectors for the custom-step variants by executing the NOTEBOOK'S OWN CODE on the numpy tf stand-in
of make_golden.py: nb-bool cell 4 (class SimpleEncoder) is exec'd verbatim, and the per-gate forward lines of its cell 6
train_step (split -> tf.random.normal -> KL) are cut out of the cell text and exec'd verbatim around it.

Run here (needs /root/reference; never on the GPU box):   python tests/golden/make_f3_golden.py
Writes tests/golden/ref_simple_encoder.npz (committed)."""
import json
import os
import sys
import textwrap

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G                                                # noqa: E402

NB = "/root/reference/complex_systems/InfoDecomp_Boolean_circuits.ipynb"


def main():
    tf = G.install_tf_shim()
    tf.ones = lambda shape, dtype=None: np.ones(shape, dtype=np.float32 if dtype in (None, "float32") else dtype)
    tf.ones_like = np.ones_like

    class _Var(G._Variable):                                            # SimpleEncoder.build: tf.Variable(initial_value=..., trainable=True)
        def __init__(self, initial_value=None, trainable=True, dtype=None):
            super().__init__(initial_value, dtype=dtype)

        def __rmul__(self, o):
            return np.asarray(o) * self._v

        __mul__ = __rmul__

    tf.Variable = _Var
    cells = ["".join(c["source"]) for c in json.load(open(NB))["cells"] if c["cell_type"] == "code"]
    cell4 = next(c for c in cells if c.lstrip().startswith("class SimpleEncoder"))
    ns = {"tf": tf, "np": np}
    exec(cell4, ns)                                                     # nb-bool cell 4, verbatim
    SimpleEncoder = ns["SimpleEncoder"]
    enc0 = SimpleEncoder()
    enc0.build(None)
    assert float(enc0.mu_scaling.value()[0, 0]) == 1.0 and float(enc0.logvar.value()[0, 0]) == -3.0   # initial values

    cell6 = next(c for c in cells if "def train_step" in c and "SimpleEncoder()" in c)
    a = cell6.index("    for gate_ind in range(number_input_gates):")
    b = cell6.index("    y_predicted = predictive_model")
    loop_src = textwrap.dedent(cell6[a:b])                              # nb-bool cell 6 train_step, the per-gate forward lines

    rng = np.random.default_rng(31)
    F, B = 10, 64
    x = (rng.integers(0, 2, size=(B, F)) * 2.0 - 1.0).astype(np.float32)     # cell 6: map 0,1 -> -1,1
    eps = rng.standard_normal((B, F, 1)).astype(np.float32)
    enc_params = np.empty(2 * F, np.float32)
    encoders = []
    for i in range(F):
        e = SimpleEncoder()
        e.build(None)
        e.mu_scaling.assign(np.full((1, 1), 0.5 + rng.random(), np.float32))
        e.logvar.assign(np.full((1, 1), -3.0 + rng.standard_normal(), np.float32))
        enc_params[2 * i], enc_params[2 * i + 1] = e.mu_scaling.value()[0, 0], e.logvar.value()[0, 0]
        encoders.append(e)
    G.EPS.q = [eps[:, i, :].astype(np.float64) for i in range(F)]
    loc = {"tf": tf, "number_input_gates": F, "feature_encoders": encoders,
           "batch_x_split": [x[:, i:i + 1].astype(np.float64) for i in range(F)], "all_embeddings": [], "kl_divergence_channels": []}
    exec(loop_src, loc)
    assert not G.EPS.q
    out = dict(x=x, eps=eps, enc_params=enc_params, emb=np.concatenate(loc["all_embeddings"], -1),
               kl=np.asarray(loc["kl_divergence_channels"], np.float64))
    for i in range(F):
        out[f"enc{i}"] = np.asarray(encoders[i](x[:, i:i + 1].astype(np.float64)))
    np.savez_compressed(os.path.join(HERE, "ref_simple_encoder.npz"), **out)
    print("simple encoder golden: kl[:3] =", out["kl"][:3], "emb[0,:3] =", out["emb"][0, :3])


if __name__ == "__main__":
    main()
