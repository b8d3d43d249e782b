"""bayesnet.model -- same module path as the reference."""
import _bootstrap  # noqa: F401
from genofix_b200.bayesnet.model import (BayesPedigreeNetworkModel, _default_alleleprobs, _default_mendelprobs,  # noqa: F401
                                         _default_mendelprobs_np, generate_probs_differences_kids,
                                         generate_probs_differences_parent, generate_probs_kids)
