"""Oracle: CPU restatement of the reference's simultaneous-sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker or as the timed CPU baseline - never on the CUDA product path.

Pinning: the reference ships no tests or golden vectors (SURVEY.md section 4,
8c), so the oracle is pinned against the reference itself, imported read-only
from /root/reference in the build container by ``tests/golden/make_golden.py``;
the resulting fixtures live in ``tests/golden/*.npz`` and
``tests/test_oracle_golden.py`` replays them on every CPU test run.

Modules (each function cites the reference file:line it follows):
  sigmas.py        geometric noise schedule       (models/__init__.py:5-18)
  weights.py       deterministic parameter sets with the reference's
                   state_dict inventory          (models/ncsnv2.py:420-477)
  scorenet_ref.py  NCSN_LiDAR_small forward, functional torch fp32
                                                  (models/ncsnv2.py:484-518)
  crossview_ref.py un-project / pose / re-project / z-buffer / fusion
                   (models/KITTISampling.py:160-430, models/__init__.py:255-520)
  samplers_ref.py  the three annealed-Langevin loops
                   (KITTISampling.py:6-513, models/__init__.py:112-602,1385-1442)
"""
