import numpy as np


class DiscreteFactor(object):
    def __init__(self, variables, cardinality, values, state_names=None):
        self.variables = list(variables)
        self.cardinality = np.array(cardinality, dtype=int)
        self.values = np.asarray(values, dtype=float).reshape(self.cardinality)
        self.state_names = state_names

    def scope(self):
        return list(self.variables)


class TabularCPD(DiscreteFactor):
    """values[k, 3*e0 + e1] for evidence=[e0, e1] (first evidence variable is the slowest axis)."""

    def __init__(self, variable, variable_card, values, evidence=None, evidence_card=None, state_names=None):
        self.variable = variable
        self.variable_card = variable_card
        evidence = list(evidence) if evidence is not None else []
        evidence_card = list(evidence_card) if evidence_card is not None else []
        self._raw_values = np.array(values, dtype=float)
        super().__init__([variable] + evidence, [variable_card] + evidence_card, self._raw_values, state_names)

    def get_evidence(self):
        return self.variables[:0:-1]

    def get_values(self):
        return self._raw_values.reshape(self.variable_card, -1)

    def copy(self):
        ev = self.variables[1:]
        return TabularCPD(self.variable, self.variable_card, self._raw_values.copy(),
                          evidence=ev if ev else None,
                          evidence_card=list(self.cardinality[1:]) if ev else None,
                          state_names=self.state_names)

    def __repr__(self):
        return "<TabularCPD %s | %s>" % (self.variable, ",".join(map(str, self.variables[1:])))
