"""CPU oracle for the per-step agent pipeline.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and only as the checker (or as the timed CPU
baseline) -- never as a fallback for the CUDA path.

PARITY UNPINNED: the reference (/root/reference/src/model.py) ships no tests, no
golden vectors and no fixtures, it cannot be imported here (``import pyflamegpu``
fails, FLAMEGPU2 has no CPU backend) and its cuRAND streams depend on FLAMEGPU2's
internal agent-list order.  This oracle is therefore our own NumPy transcription
of the reference's agent functions, keyed with counter-based Philox; the only
external pins are the Random123 Philox known-answer vectors and hand-derived
micro-scenarios (tests/test_oracle_*.py).
"""
