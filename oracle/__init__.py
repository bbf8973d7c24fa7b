"""CPU oracle of the MC-scoring hot path — TEST INFRASTRUCTURE ONLY.

Nothing under `oracle/` is part of the product: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s CPU-baseline / `--impl reference` legs may import it, and only as the checker or as
the thing timed *as the CPU baseline*.  The product path (bayesformer_b200/) never imports it.

Parity status: PINNED against the reference itself.  The reference has no tests or golden vectors
for this path; the restatement in `restatement.py` is checked against outputs of the
unmodified reference run in the build container (`oracle/make_golden.py` → `tests/golden/*.npz`,
masks recorded from the reference's own `F.dropout` calls) and against the known-answer vectors
of SURVEY §4.
"""
