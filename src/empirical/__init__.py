"""Reference module path kept for drop-in use (PYTHONPATH=src); implementation in genofix_b200."""
