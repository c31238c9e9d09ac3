"""Host mirror of the reference's ``ctw`` extension module (chaos/ctw.pyx:2-3, built by chaos/setup.py):
``ctw.estimate_entropy(seq, alphabet_size)`` -> bits/symbol, through the C ABI (dib_ctw_estimate_entropy*).
These are HOST functions -- the context-tree build stays on the CPU by design."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


def _symbols(seq):
    a = np.asarray(seq)
    if a.ndim != 1:
        raise ValueError("sequence must be one-dimensional")
    if a.size and (a.min() < 0 or a.max() > 126):
        raise ValueError("symbols must lie in [0, 127)")
    return np.ascontiguousarray(a, dtype=np.int8)


def estimate_entropy(seq, alphabet_size):
    """ctw.estimate_entropy (chaos/ctw.pyx:2-3 -> cppctw.cpp:163-171)."""
    lib = _lib.load()
    s = _symbols(seq)
    out = ctypes.c_double()
    if lib.dib_ctw_estimate_entropy(ctypes.c_void_p(s.ctypes.data), s.size, int(alphabet_size), ctypes.byref(out)) != 0:
        raise _lib.DibError(lib.dib_ctw_last_error().decode())
    return out.value


def estimate_entropy_batch(sequences, alphabet_size, num_threads=0):
    """Independent sequences (e.g. the 75 sub-sequences of one partition evaluation, nb-chaos cell 3) on a host thread
    pool.  Returns a float64 array, one estimate per sequence."""
    lib = _lib.load()
    seqs = [_symbols(s) for s in sequences]
    offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([s.size for s in seqs], out=offsets[1:])
    flat = np.concatenate(seqs) if seqs else np.zeros(0, dtype=np.int8)
    out = np.empty(len(seqs), dtype=np.float64)
    rc = lib.dib_ctw_estimate_entropy_batch(ctypes.c_void_p(flat.ctypes.data), ctypes.c_void_p(offsets.ctypes.data), len(seqs),
                                            int(alphabet_size), int(num_threads), ctypes.c_void_p(out.ctypes.data))
    if rc != 0:
        raise _lib.DibError(lib.dib_ctw_last_error().decode())
    return out
