"""Checker-side probes (they import `oracle/`, which only `tests/`, `smoke()` and the CPU legs of `bench.py` may do): precision of the
reduced-precision evaluator paths and of a would-be Winograd tower against the reference-generated goldens.  Not collected by pytest."""
