"""Mendel model with the reference's call surface (bayesnet/model.py).

The probability tables are the reference's constants (model.py:24-30).  The children heuristic and
the blanket inference run in the native library (the same routines the kernels use):
gfx_kids_term_host restates generate_probs_kids (model.py:65-89) including numpy's summation order,
gfx_infer_host replaces the pgmpy query on the network generateModel builds (model.py:173-230)."""
import ctypes as C

import numpy as np

from .. import _lib

_default_alleleprobs = [[0.25], [0.5], [0.25]]
#                          sire 0  0   0   1   1    1  2   2  2
#                           dam 0  1   2   0   1    2  0   1  2
_default_mendelprobs = [[1, .5, 0, .5, .25, 0, 0, 0, 0],
                        [0, .5, 1, .5, .5, .5, 1, .5, 0],
                        [0, 0, 0, 0, .25, .5, 0, .5, 1]]
_default_mendelprobs_np = np.array(_default_mendelprobs)
_T3 = _default_mendelprobs_np.reshape(3, 3, 3)
_THIRD16 = np.float16(1365 / 4096)   # float16(1/3), the unit of generate_probs_differences_kids


def _codes(values):
    return np.array([int(v) if (v is not None and v in (0, 1, 2)) else 3 for v in values], dtype=np.uint8)


def generate_probs_kids(kidstates, oneparentstates):
    """P(parent state) suggested by the children (model.py:65-89)."""
    kid = _codes(kidstates)
    par = _codes(oneparentstates)
    if len(kid) == 0 or not (kid <= 2).any():
        return [1 / 3, 1 / 3, 1 / 3]
    out = np.zeros(3)
    rc = _lib.lib().gfx_kids_term_host(len(kid), kid.ctypes.data_as(C.c_void_p), par.ctypes.data_as(C.c_void_p),
                                       out.ctypes.data_as(C.c_void_p))
    _lib.check(rc)
    return out


def generate_probs_differences_kids(kidstates, oneparentstates, observedstate):
    """Per usable child: max - value[observed] of the child's float16 parent-state vector
    (model.py:93-116).  Only four vectors occur."""
    kid = _codes(kidstates)
    par = _codes(oneparentstates)
    out = []
    if not (kid <= 2).any():
        return out
    z = np.float16(0)
    A = [np.float16(2) * _THIRD16, _THIRD16, z]   # [2/3, 1/3, 0]
    B = [_THIRD16, _THIRD16, _THIRD16]
    Cc = [z, _THIRD16, np.float16(2) * _THIRD16]
    Z = [z, z, z]
    for k, p in zip(kid, par):
        if k > 2:
            continue
        if k == 0:
            v = Z if p == 2 else A
        elif k == 2:
            v = Z if p == 0 else Cc
        else:
            v = Cc if p == 0 else (A if p == 2 else B)
        out.append(np.float16(max(v)) - v[observedstate])
    # the reference returns one entry per element of kidstates: the rows of the usable children first, then the
    # untouched zero rows of its float16 table (model.py:100,113-115)
    out.extend([np.float16(0)] * int((kid > 2).sum()))
    return out


def generate_probs_differences_parent(parent1, parent2, observedstate):
    """max - value[observed] of the Mendel column selected by the parents (model.py:118-133)."""
    k1, k2 = parent1 in (0, 1, 2), parent2 in (0, 1, 2)
    if k1 and k2:
        column = _T3[:, parent1, parent2]
    elif k1:
        column = np.nanmean(_T3[:, parent1, :], 1, dtype=float)
    elif k2:
        column = np.nanmean(_T3[:, :, parent2], 1, dtype=float)
    else:
        return 0
    return np.nanmax(column) - column[observedstate]


class BayesPedigreeNetworkModel(object):
    """The network of a (sub-)pedigree: founders of the subset get the allele prior, every other node
    the Mendel table (model.py:173-230).  Nodes are kept in topological order for gfx_infer_host."""

    def __init__(self, nodes=(), parents=None):
        self._nodes = [str(n) for n in nodes]
        self._parents = dict(parents or {})
        self.blanket = None
        self.latents = set()

    @property
    def nodes(self):
        return list(self._nodes)

    def edges(self):
        return [(p, k) for k, ps in self._parents.items() for p in ps]

    def copy(self):
        m = BayesPedigreeNetworkModel(self._nodes, self._parents)
        m.blanket = self.blanket
        return m

    def check_model(self):
        return True

    @staticmethod
    def generateModel(pedigree, alleleprobs=_default_alleleprobs, mendelprobs=_default_mendelprobs):
        if alleleprobs != _default_alleleprobs or mendelprobs != _default_mendelprobs:
            raise NotImplementedError("only the default allele prior / Mendel table are supported on the device path")
        k2s, k2d = pedigree.kid2sire, pedigree.kid2dam
        parents = {}
        for kid, sire in k2s.items():
            parents[str(kid)] = (str(sire), str(k2d[kid]))   # KeyError for one known parent, like the reference
        nodes = set(parents)
        for s, d in parents.values():
            nodes.update((s, d))
        order, done = [], set()

        def visit(n):
            stack = [n]
            while stack:
                x = stack[-1]
                pend = [p for p in parents.get(x, ()) if p not in done]
                if x in done:
                    stack.pop()
                elif pend:
                    stack.extend(pend)
                else:
                    done.add(x)
                    order.append(x)
                    stack.pop()
        for n in sorted(nodes):
            visit(n)
        return BayesPedigreeNetworkModel(order, parents)

    def query(self, variables, evidence=None):
        """Exact marginals {var: ndarray(3)} given hard evidence {node: state}."""
        evidence = {str(k): int(v) for k, v in (evidence or {}).items()}
        n = len(self._nodes)
        pos = {v: i for i, v in enumerate(self._nodes)}
        p1 = np.full(n, -1, np.int8)
        p2 = np.full(n, -1, np.int8)
        code = np.full(n, 3, np.uint8)
        for k, (s, d) in self._parents.items():
            p1[pos[k]], p2[pos[k]] = pos[s], pos[d]
        for k, v in evidence.items():
            if k in pos and v in (0, 1, 2):
                code[pos[k]] = v
        out = {}
        L = _lib.lib()
        for q in variables:
            c = code.copy()
            for other in variables:
                c[pos[str(other)]] = 3
            res = np.zeros(3)
            _lib.check(L.gfx_infer_host(n, p1.ctypes.data_as(C.c_void_p), p2.ctypes.data_as(C.c_void_p),
                                        c.ctypes.data_as(C.c_void_p), pos[str(q)], res.ctypes.data_as(C.c_void_p)))
            out[str(q)] = res
        return out
