import networkx as nx


class DAG(nx.DiGraph):
    # pgmpy.base.DAG.__init__(ebunch=None, latents=set())
    def __init__(self, ebunch=None, latents=set()):
        super().__init__(ebunch)
        self.latents = set(latents)


class BayesianModel(DAG):
    def __init__(self, ebunch=None, latents=set()):
        super().__init__(ebunch=ebunch, latents=latents)
        self.cpds = []

    def add_cpds(self, *cpds):
        for cpd in cpds:
            if cpd.variable not in self.nodes:
                raise ValueError("CPD defined on variable not in the model", cpd)
            for i, old in enumerate(self.cpds):
                if old.variable == cpd.variable:
                    self.cpds[i] = cpd
                    break
            else:
                self.cpds.append(cpd)

    def get_cpds(self, node=None):
        if node is None:
            return self.cpds
        for cpd in self.cpds:
            if cpd.variable == node:
                return cpd
        return None

    def get_parents(self, node):
        return list(self.predecessors(node))

    def check_model(self):
        for node in self.nodes():
            cpd = self.get_cpds(node)
            if cpd is None:
                raise ValueError("No CPD associated with {n}".format(n=node))
            if set(cpd.get_evidence()) != set(self.get_parents(node)):
                raise ValueError("CPD associated with {n} doesn't have proper parents associated with it.".format(n=node))
        return True


BayesianNetwork = BayesianModel
