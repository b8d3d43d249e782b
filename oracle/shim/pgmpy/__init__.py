"""Minimal stand-in for pgmpy 0.1.16 (TEST INFRASTRUCTURE ONLY).

pgmpy is pinned by /root/reference/README.md:9 but is not installed in this image and
cannot be fetched (no network).  This package exposes exactly the names the reference
imports (src/bayesnet/model.py:12-17, src/correct_genotypes3.py:12-13) so that the
UNMODIFIED reference can be imported and executed to generate golden vectors.

The only arithmetic restated here is exact discrete inference
(`pgmpy.inference.VariableElimination.query`); everything else is a no-op container.
Never imported by the product (genofix_b200/).
"""
__version__ = "0.1.16-shim"
