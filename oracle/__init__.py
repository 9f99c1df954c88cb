"""CPU oracle — test infrastructure only (see numpy_oracle.py / stdbscan_ref.c headers).

Importable from ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
reference arm; never from the product package.
"""
