"""oracle/ -- TEST INFRASTRUCTURE ONLY.

A CPU (plain torch, fp32 or fp64) restatement of the radar_depth hot path, written
as pure functions over a reference-format ``state_dict``.  It exists to CHECK the
CUDA product in ``radar_depth_b200``; nothing in the product imports it.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.

Parity status: **pinned** -- ``tests/golden/*.npz`` hold outputs of the real reference
modules (``/root/reference/model/models.py``, ``model/multistage_model.py``,
``evaluation/criteria_new.py``) generated in the build container by
``oracle/gen_golden.py``; ``tests/test_oracle_golden.py`` checks this restatement
against them, and (when ``/root/reference`` is present) against the live reference.
The reference repo itself ships no tests or golden vectors.
"""
