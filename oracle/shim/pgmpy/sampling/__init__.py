class BayesianModelSampling(object):
    pass
