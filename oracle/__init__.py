"""Oracle package — TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the shipped product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker / the timed CPU baseline.
"""
