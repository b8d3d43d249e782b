"""Host logic of the reference's static call surface (initializer + CorrectGenotypes.mendelProbsSingle / correctPair /
correctSingle in genofix_b200/correct_genotypes3.py) on CPU: the engine is replaced by a stand-in whose "kernels" are the
oracle, so what is checked is the product's Python around them -- the identity rank table that feeds caller-supplied E
values to the kernels, the mapping of schedule jobs to (sire, dam) pairs, the assembly of ResultPairInfer /
ResultSingleInfer and the lone-node / pair-member conventions.  The same scenario runs against the real kernels in
tests/test_gpu_parity.py::test_static_call_surface_matches_oracle."""
import numpy as np
import pytest
import torch

from genofix_b200 import _lib, synth
from genofix_b200.engine import Schedule
from oracle.genofix_oracle import Oracle


class _OracleKernels(object):
    """gfx_build_cut_table is the real (host-only) entry point; the two correction entry points are the oracle."""

    def __init__(self, st):
        self.st = st
        self.real = _lib.lib()

    def gfx_build_cut_table(self, *a):
        return self.real.gfx_build_cut_table(*a)

    def gfx_correct_pairs(self, handle, g, live, s, wpr, raw, ld_raw, elut, pp, snap, snapn, post, tall, stream):
        st = self.st
        eng = st["eng"]
        seen = eng.host_elut.numpy()[eng.raw[:, :eng.s].numpy().view(np.uint16)]
        assert np.array_equal(seen, st["E"], equal_nan=True)      # the identity table reproduces the caller's E
        gen_off, jobs = eng.schedule.array(1), eng.schedule.array(3)
        pairs = eng.schedule.array(0).reshape(-1, 2)
        for k in range(int(gen_off[g + 1] - gen_off[g])):
            rs, rd = pairs[jobs[gen_off[g] + k]]
            res = st["o"].correct_pair(st["obs"], st["E"], st["ids"][rs], st["ids"][rd], 3, st["test_p"])
            for j, (ps, pd_, _a, _b) in (res or {}).items():
                st["post"][2 * k, j] = torch.from_numpy(ps)
                st["post"][2 * k + 1, j] = torch.from_numpy(pd_)
        return 0

    def gfx_correct_singles(self, handle, live, s, wpr, raw, ld_raw, elut, pp, snap, snapn, post, tall, stream):
        st = self.st
        for k, r in enumerate(st["eng"].schedule.array(2)):
            res = st["o"].correct_single(st["obs"], st["E"], st["ids"][r], 2, st["test_p"])
            for j, (p, _op) in (res or {}).items():
                st["post"][k, j] = torch.from_numpy(p)
        return 0


class _StandInEngine(object):
    def __init__(self, st, ped, ids, back):
        self.st, self.torch, self.L = st, torch, _OracleKernels(st)
        self.schedule = Schedule(ped, ids, back, host_only=True)
        self.n, self.back = len(ids), back
        st["eng"] = self

    def _alloc(self, s):
        self.s = s
        self.wpr = (((s + 15) // 16 + 3) // 4) * 4
        self.ld_raw = 16 * self.wpr
        n, ldc = self.n, ((s + 15) // 16) * 16
        self.codes = torch.zeros((n, ldc), dtype=torch.uint8)
        self.host_in = torch.zeros((n, ldc), dtype=torch.uint8)
        self.packed_orig = torch.zeros((n, self.wpr), dtype=torch.int32)
        self.packed_live = torch.zeros((n, self.wpr), dtype=torch.int32)
        self.raw = torch.zeros((n, self.ld_raw), dtype=torch.int16)
        self.host_elut, self.elut = torch.zeros(65536, dtype=torch.float64), torch.zeros(65536, dtype=torch.float64)
        self.host_cutraw, self.cutraw = torch.zeros(65537, dtype=torch.int16), torch.zeros(65537, dtype=torch.int16)
        self.snapshot = torch.zeros(16, dtype=torch.int32)

    def load_codes_device(self, codes):
        pass

    def mendel_rank_device(self, fused=True):
        r = self.st["o"].mendel_rank(self.st["obs"], self.back).view(np.uint16).copy()
        r[self.st["obs"] == 9] = 0xFFFF
        self.raw[:, :self.s] = torch.from_numpy(r.view(np.int16))

    def check_status(self):
        pass

    def _stream(self):
        return None

    def raw_host(self):
        r = self.raw[:, :self.s].numpy().view(np.uint16).copy()
        r[r == 0xFFFF] = 0x3C00
        return r.view(np.float16)


def test_static_call_surface_host_logic(monkeypatch):
    import genofix_b200.correct_genotypes3 as cg
    import test_gpu_parity as gpu_tests
    rows = synth.make_pedigree(150, 5, sire_frac=0.25, dam_frac=0.5, seed=21, related_mating_p=0.3)
    obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, 40, seed=3), 0.04, seed=4), 0.03, seed=5)
    ids = [int(x) for x in rows[:, 0]]
    o = Oracle(rows, ids)
    E, test_p = o.transform(o.mendel_rank(obs, 2), obs, 0.9)
    st = dict(o=o, obs=obs, ids=ids, E=E, test_p=test_p)
    real_full = torch.full

    def full(shape, val, dtype=None, device=None):       # the statics allocate the posterior buffer with torch.full
        st["post"] = real_full(shape, val, dtype=dtype)
        return st["post"]
    monkeypatch.setattr(torch, "full", full)
    monkeypatch.setattr(cg, "engine_for", lambda ped, ids_, back: _StandInEngine(st, ped, ids_, back))
    scenario = gpu_tests.test_static_call_surface_matches_oracle
    getattr(scenario, "__wrapped__", scenario)()
