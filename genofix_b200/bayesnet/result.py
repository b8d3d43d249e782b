"""Result records returned by correctPair / correctSingle (reference: bayesnet/result.py:11-32,47-61).

Plain per-SNP dict containers, keyed by SNP id, so that callers written against the reference keep
working; the GPU path fills them from the posterior buffers of gfx_correct_pairs/_singles."""


class ResultPairInfer(object):
    def __init__(self, sire, dam, model):
        self.sire = sire
        self.dam = dam
        self.model = model
        for name in ("jointprobs", "bestprobs", "beststate", "bestprob_sire", "bestprob_dam",
                     "prob_sire", "prob_dam", "observedprob_sire", "observedprob_dam"):
            setattr(self, name, {})


class ResultSingleInfer(object):
    def __init__(self, kid, model):
        self.kid = kid
        self.model = model
        for name in ("bestprobs", "beststate", "observedprob", "prob"):
            setattr(self, name, {})


class ResultGroupInfer(object):
    def __init__(self, sire, dam, model):
        self.results = []
        self.model = model
