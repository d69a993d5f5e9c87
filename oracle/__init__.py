"""CPU oracle for the PSRS hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it,
and there only as the checker / the timed CPU baseline, never as a fallback.

Parity status: PINNED.  The restatement is checked against (i) NumPy itself
for the RNG layer (``tests/test_oracle_rng.py``) and (ii) golden vectors made
by running the UNMODIFIED reference (``/root/reference/offsim4rl``) in the
build container — ``tests/golden/make_golden.py`` is the generating script,
``tests/golden/*.npz`` the committed fixtures.
"""
