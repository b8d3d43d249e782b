class BayesianEstimator(object):
    pass


class MaximumLikelihoodEstimator(object):
    pass
