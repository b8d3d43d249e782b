"""correct_genotypes2 -- module path of the reference's earlier driver.  pedigree/error_checker.py:164-187 and
empirical/preindex.py:230-253 import it for `initializer` and `CorrectGenotypes.mendelProbsSingle`, the twin of
correct_genotypes3's (correct_genotypes2.py:62-150; same scores, checked through the error_checker fixtures).  Both
are the CUDA mirror's.  v2's own correctMatrix (in-run empirical clusters, :515) is not part of the path."""
import _bootstrap  # noqa: F401
from genofix_b200.correct_genotypes3 import CorrectGenotypes as _V3, initializer, initializerEmp  # noqa: F401


class CorrectGenotypes(_V3):
    def correctMatrix(self, *args, **kwargs):
        raise NotImplementedError("correct_genotypes2.correctMatrix is not mirrored; use correct_genotypes3.CorrectGenotypes")
