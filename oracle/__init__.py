"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference scoring path.

Nothing under ``oracle/`` may be imported by the product package
(``multi-modal_colpali_b200/``).  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` -- and there only as the checker / the reported CPU
baseline, never as the thing shipped.
"""
