"""CPU oracle for the mastering DSP chain.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker
(or as the timed CPU arm) - never as a fallback for the CUDA path.  See ``oracle/README.md``.
"""
