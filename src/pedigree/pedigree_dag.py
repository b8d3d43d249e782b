"""pedigree.pedigree_dag -- same module path as the reference; flat-array implementation."""
import _bootstrap  # noqa: F401
from genofix_b200.pedigree.pedigree_dag import PedigreeDAG  # noqa: F401
