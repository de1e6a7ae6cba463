"""ORACLE — test infrastructure only (see oracle/torch_ref.py and oracle/c/*.c headers).

Nothing under `cddmsl_b200/` imports this package; only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs do.
"""
