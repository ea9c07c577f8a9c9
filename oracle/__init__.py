"""CPU oracle for the ISBFSAR `modules/ar` scoring path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``isbfsar_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker
or as the timed CPU baseline -- never as the product path.

Parity status: PINNED.  ``oracle/gen_golden.py`` imports the unmodified
reference (``/root/reference/modules/ar/utils/model.py`` and
``modules/hpe/utils/misc.py``) in the build container, runs it on the
deterministic synthetic weights/inputs of ``oracle/synth.py`` and freezes the
outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this
restatement against those vectors.
"""
