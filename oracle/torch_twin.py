"""TEST / BASELINE INFRASTRUCTURE ONLY -- PyTorch-CPU eager twin of the reference graph.

TensorFlow is unavailable here (see dib_oracle.py header), so the reference's tf.keras CPU path cannot be timed.
This twin mirrors its STRUCTURE op for op -- a Python loop over F independent Sequential encoders, one RNG call
and one KL reduction per feature, concat, integration MLP, autograd, one Adam update per variable with Keras
semantics (models.py:96-123; Model.fit train_step) -- and is what ``bench.py --impl reference`` and the
``cpu_baseline`` leg time on the GPU box's host cores.  It is also used by tests to cross-check the numpy
oracle's hand-written backward against autograd.  The product path never imports it.
"""
import math

import numpy as np
import torch

from . import dib_oracle as O

_ACT = {None: lambda: torch.nn.Identity(), "linear": lambda: torch.nn.Identity(), "relu": lambda: torch.nn.ReLU(),
        "tanh": lambda: torch.nn.Tanh(), "sigmoid": lambda: torch.nn.Sigmoid(), "elu": lambda: torch.nn.ELU(),
        "leaky_relu": lambda: torch.nn.LeakyReLU(0.2)}


class PositionalEncoding(torch.nn.Module):           # models.py:12-23
    def __init__(self, frequencies):
        super().__init__()
        self.frequencies = frequencies

    def forward(self, x):
        return torch.cat([x] + [torch.sin(f * x) for f in self.frequencies], -1)


class TwinDIB(torch.nn.Module):
    def __init__(self, cfg: O.DIBConfig, flat_params=None):
        super().__init__()
        self.cfg = cfg
        encs = []
        for i in range(cfg.number_features):         # models.py:71-78
            dims = cfg.encoder_layer_dims(i)
            layers = [PositionalEncoding(cfg.frequencies)] if cfg.use_positional_encoding else []
            for k in range(len(dims) - 1):
                layers.append(torch.nn.Linear(dims[k], dims[k + 1]))
                if k < len(dims) - 2:
                    layers.append(_ACT[cfg.activation_fn]())
            encs.append(torch.nn.Sequential(*layers))
        self.feature_encoders = torch.nn.ModuleList(encs)
        dims = cfg.integration_layer_dims()          # models.py:81-84
        layers = []
        for k in range(len(dims) - 1):
            layers.append(torch.nn.Linear(dims[k], dims[k + 1]))
            layers.append(_ACT[cfg.activation_fn if k < len(dims) - 2 else cfg.output_activation_fn]())
        self.integration_network = torch.nn.Sequential(*layers)
        self.beta = 1.0
        if flat_params is not None:
            self.load_flat(flat_params)

    def linears(self):
        out = []
        for enc in self.feature_encoders:
            out += [m for m in enc if isinstance(m, torch.nn.Linear)]
        out += [m for m in self.integration_network if isinstance(m, torch.nn.Linear)]
        return out

    def load_flat(self, flat):
        flat = np.asarray(flat)
        off = 0
        with torch.no_grad():
            for lin in self.linears():
                fi, fo = lin.in_features, lin.out_features
                lin.weight.copy_(torch.from_numpy(flat[off:off + fi * fo].reshape(fi, fo).T.copy()).to(lin.weight.dtype))
                off += fi * fo
                lin.bias.copy_(torch.from_numpy(flat[off:off + fo].copy()).to(lin.bias.dtype))
                off += fo
        assert off == flat.size

    def flat_grads(self):
        out = []
        for lin in self.linears():
            out += [lin.weight.grad.T.reshape(-1), lin.bias.grad.reshape(-1)]
        return torch.cat(out)

    def forward(self, x, eps=None):
        cfg = self.cfg
        E = cfg.feature_embedding_dimension
        feats = torch.split(x, list(cfg.feature_dimensionalities), dim=-1)        # models.py:101
        embs, kls = [], []
        for i, enc in enumerate(self.feature_encoders):                           # models.py:105
            mu, lv = torch.split(enc(feats[i]), E, dim=-1)
            noise = torch.randn_like(mu) if eps is None else eps[:, i, :]         # models.py:108
            embs.append(mu + torch.exp(lv / 2.) * noise)
            kls.append(torch.mean(torch.sum(0.5 * (mu ** 2 + torch.exp(lv) - lv - 1.), dim=-1)))   # :111-112
        pred = self.integration_network(torch.cat(embs, -1))                      # models.py:122
        return pred, torch.stack(kls)

    def loss(self, x, y, loss_kind, eps=None):
        pred, kls = self(x, eps)
        if loss_kind == O.LOSS_BCE_LOGITS:
            task = torch.nn.functional.binary_cross_entropy_with_logits(pred, y.reshape(pred.shape))
        elif loss_kind == O.LOSS_SPARSE_CE_LOGITS:
            task = torch.nn.functional.cross_entropy(pred, y.reshape(-1).long())
        else:
            task = torch.nn.functional.mse_loss(pred, y.reshape(pred.shape))
        return task + self.beta * kls.sum(), task, kls, pred


class KerasAdam:
    """One update per variable, epsilon outside the bias correction (Keras)."""

    def __init__(self, params, lr, b1=0.9, b2=0.999, eps=1e-7):
        self.params = list(params)
        self.lr, self.b1, self.b2, self.eps, self.t = lr, b1, b2, eps, 0
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]

    @torch.no_grad()
    def step(self):
        self.t += 1
        lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        for p, m, v in zip(self.params, self.m, self.v):
            g = p.grad
            m.add_(g - m, alpha=1 - self.b1)
            v.add_(g * g - v, alpha=1 - self.b2)
            p.addcdiv_(m, v.sqrt() + self.eps, value=-lr_t)
            p.grad = None


def time_train_steps(cfg, loss_kind, x, y, lr, steps, warmup, threads=None):
    """Seconds per training step (median over ``steps``) of the eager twin on the host cores."""
    import time
    if threads:
        torch.set_num_threads(threads)
    model = TwinDIB(cfg)
    opt = KerasAdam(model.parameters(), lr)
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        total, _, _, _ = model.loss(xt, yt, loss_kind)
        total.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return float(np.median(times)), float(np.sum(times)), torch.get_num_threads()
