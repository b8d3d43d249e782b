"""mgzip -> gzip alias (TEST INFRASTRUCTURE ONLY)."""
from gzip import *  # noqa
from gzip import open  # noqa
