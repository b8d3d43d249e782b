"""mgzip -> gzip alias (TEST INFRASTRUCTURE ONLY).  mgzip.open takes `thread` / `blocksize` (gfutils/pickle_util.py:16)."""
import gzip as _gzip
from gzip import *  # noqa


def open(filename, mode="rb", compresslevel=9, encoding=None, errors=None, newline=None, thread=None, blocksize=None):  # noqa: A001
    return _gzip.open(filename, mode, compresslevel, encoding, errors, newline)
