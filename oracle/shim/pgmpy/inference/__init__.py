"""Exact inference used in place of pgmpy 0.1.16 `VariableElimination.query`.

Semantics restated (from the published pgmpy algorithm; pgmpy source is NOT on this machine):
  * factors are the CPDs, reduced by the hard evidence;
  * a factor whose scope becomes empty after reduction / elimination is dropped (pgmpy never
    re-registers a factor with no variables), so contradictory evidence that is d-separated
    from the query does not poison the result;
  * non-query variables are summed out (min-neighbours order here; order only changes rounding,
    and all unnormalised quantities on this path are exact dyadic rationals);
  * joint=False: product of the remaining factors, marginalise to each query variable,
    divide by the sum (0/0 -> NaN, as `DiscreteFactor.normalize`).
"""
import numpy as np


class _Marginal(object):
    def __init__(self, var, values):
        self.variables = [var]
        self.values = values


def _multiply(f1, f2):
    v1, t1 = f1
    v2, t2 = f2
    allv = list(v1) + [v for v in v2 if v not in v1]
    idx = {v: i for i, v in enumerate(allv)}
    out = np.einsum(t1, [idx[v] for v in v1], t2, [idx[v] for v in v2], [idx[v] for v in allv])
    return (allv, out)


class VariableElimination(object):
    def __init__(self, model):
        self.model = model

    def query(self, variables, evidence=None, virtual_evidence=None, elimination_order="MinFill",
              joint=True, show_progress=True):
        evidence = {} if evidence is None else {k: int(v) for k, v in evidence.items()}
        variables = list(variables)
        if set(variables) & set(evidence):
            raise ValueError("query variable in evidence")
        factors = []
        for cpd in self.model.cpds:
            vs = list(cpd.variables)
            t = np.asarray(cpd.values, dtype=float)
            keep = []
            index = []
            for v in vs:
                if v in evidence:
                    index.append(evidence[v])
                else:
                    index.append(slice(None))
                    keep.append(v)
            t = t[tuple(index)]
            if len(keep) == 0:
                continue  # scalar factor: dropped
            factors.append((keep, t))
        # connected component of the query variables only matters through dropping of scalars
        to_eliminate = set(v for vs, _ in factors for v in vs) - set(variables)
        while to_eliminate:
            def nbrs(x):
                s = set()
                for vs, _ in factors:
                    if x in vs:
                        s.update(vs)
                s.discard(x)
                return len(s)
            var = min(sorted(to_eliminate), key=nbrs)
            inv = [f for f in factors if var in f[0]]
            factors = [f for f in factors if var not in f[0]]
            prod = inv[0]
            for f in inv[1:]:
                prod = _multiply(prod, f)
            vs, t = prod
            ax = vs.index(var)
            t = t.sum(axis=ax)
            vs = [v for v in vs if v != var]
            if len(vs) > 0:
                factors.append((vs, t))
            to_eliminate.discard(var)
        if len(factors) == 0:
            prod = ([], np.array(1.0))
        else:
            prod = factors[0]
            for f in factors[1:]:
                prod = _multiply(prod, f)
        vs, t = prod
        # any query variable that lost all its factors (cannot happen: own CPD always has it)
        if joint:
            raise NotImplementedError("shim only implements joint=False")
        out = {}
        for q in variables:
            axes = tuple(i for i, v in enumerate(vs) if v != q)
            m = t.sum(axis=axes) if axes else t
            with np.errstate(invalid="ignore", divide="ignore"):
                m = m / m.sum()
            out[q] = _Marginal(q, np.asarray(m, dtype=float))
        return out


class BeliefPropagation(VariableElimination):
    pass
