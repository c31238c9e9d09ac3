"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference's Context-Tree-Weighting entropy-rate estimator
(chaos/cppctw.cpp; ctw.estimate_entropy, chaos/ctw.pyx:2-3).  Plain Python over dict nodes: for SMALL sequences only.
Only tests/ may import this.  Pinned by oracle/_ref/libctw_ref.so (the reference's own sources compiled by oracle/Makefile)
through tests/golden/ref_ctw.npz and by the SURVEY section 4 known answers; agreement is bit-exact (same libm)."""
import ctypes
import math
import os

import numpy as np

MAX_DEPTH = 512                                   # cppctw.cpp:13


class _Node:
    __slots__ = ("counts", "children", "tail_ind", "tail_symbol", "weighted")

    def __init__(self, alphabet_size, tail_ind=-1, tail_symbol=-1):   # cppctw.cpp:27-39
        self.counts = [0] * alphabet_size
        self.children = [None] * alphabet_size
        self.tail_ind, self.tail_symbol = tail_ind, tail_symbol


def _process_sequence(seq, A):
    """cppctw.cpp:106-154: for every position walk the context backwards, counting the symbol at every node met;
    leaves are lazy 'tails' (position, symbol) pushed one level deeper each time they are walked through."""
    root = _Node(A)
    for pos, cur in enumerate(seq):
        node = root
        node.counts[cur] += 1
        for c in range(pos - 1, -1, -1):
            if node.tail_ind > 0:                                     # :121-129
                back = seq[node.tail_ind - 1]
                node.children[back] = _Node(A, node.tail_ind - 1, node.tail_symbol)
                node.children[back].counts[node.tail_symbol] += 1
                node.tail_ind, node.tail_symbol = -1, -1
            ctx = seq[c]
            if node.children[ctx] is None:                            # :132-146
                if pos - c > MAX_DEPTH:
                    break
                node.children[ctx] = _Node(A, c, cur) if c > 0 else _Node(A)
                node.children[ctx].counts[cur] += 1
                break
            node = node.children[ctx]
            node.counts[cur] += 1
    return root


def _code_lengths(root, A):
    """cppctw.cpp:57-81, post-order without recursion: KT estimate with beta = 1/A, then the CTW mixture."""
    beta = 1.0 / A
    order, stack = [], [root]
    while stack:
        n = stack.pop()
        order.append(n)
        stack.extend(ch for ch in n.children if ch is not None)
    for n in reversed(order):
        total = float(sum(n.counts))
        le = math.lgamma(total + A * beta) - math.lgamma(A * beta)
        for c in n.counts:
            le -= math.lgamma(c + beta) - math.lgamma(beta)
        le /= math.log(2)
        kids = [ch for ch in n.children if ch is not None]
        lc = 0.0
        for ch in kids:
            lc += ch.weighted
        if kids and total > 1:
            n.weighted = 1 + min(lc, le) - math.log2(1 + math.pow(2, -abs(le - lc)))
        else:
            n.weighted = le
    return root.weighted


def estimate_entropy(sequence, alphabet_size):
    """bits/symbol; the reference returns it through a float (cppctw.cpp:100-104)."""
    seq = [int(s) for s in sequence]
    if not seq:
        return float("nan")
    root = _process_sequence(seq, int(alphabet_size))
    return float(np.float32(_code_lengths(root, int(alphabet_size)) / len(seq)))


# ---- the reference's own build (oracle/_ref, see oracle/Makefile) -------------------------------------------------
_REF_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libctw_ref.so")


def reference_available():
    return os.path.exists(_REF_SO)


def reference_estimate_entropy(sequence, alphabet_size):
    lib = ctypes.CDLL(_REF_SO)
    lib.ctw_ref_estimate_entropy.restype = ctypes.c_double
    lib.ctw_ref_estimate_entropy.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int]
    s = np.ascontiguousarray(sequence, dtype=np.int8)
    return float(lib.ctw_ref_estimate_entropy(s.ctypes.data, len(s), int(alphabet_size)))
