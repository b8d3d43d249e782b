"""correct_genotypes3 -- same module path as the reference; CUDA implementation."""
import _bootstrap  # noqa: F401
from genofix_b200.correct_genotypes3 import CorrectGenotypes, initializer, initializerEmp  # noqa: F401
