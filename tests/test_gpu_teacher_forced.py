"""GPU: teacher-forced parity along the oracle's trajectory over the whole 232-level schedule with the RAW score network
(tests/tools/teacher_forced_sweep.py; the full V = 8 sweep with per-level numbers is committed under profiles/).  Here a
smaller group walks the same schedule and a spread of levels is measured: at each of them the cross-view block on
identical input must be bit-exact in every integer, the score within the arm's tolerance, and the cells whose candidate
count changes because the score differs in its last digits (flips) stay a small, reported fraction."""
import pytest

from tests.tools.teacher_forced_sweep import TOL, sweep

pytestmark = pytest.mark.gpu
LEVELS = {0, 1, 2, 3, 20, 60, 116, 180, 231}


@pytest.mark.parametrize("precision,views,max_flip_frac,x_tol", [("bf16x3", 4, 2e-3, 1e-3), ("bf16", 2, 5e-2, 5e-2)])
def test_teacher_forced_levels(precision, views, max_flip_frac, x_tol):
    res = sweep(precision, V=views, levels=LEVELS, verbose=False)
    s = res["summary"]
    print(f"[teacher-forced {precision} V={views}] score max {s['score_rel_max']:.2e} x max {s['x_rel_max']:.2e} "
          f"flipped cells max {s['flipped_cells_max']} ({s['flipped_frac_max']:.2e} of the filled cells), "
          f"exact integer steps {s['exact_integer_steps']}/{s['shared_steps']}, {s['seconds']:.0f} s")
    assert s["steps_measured"] == len(LEVELS)
    assert s["exact_integer_steps"] == s["shared_steps"] == len([c for c in LEVELS if c >= 2])
    assert s["score_rel_max"] <= TOL[precision]
    assert s["flipped_frac_max"] <= max_flip_frac
    assert s["x_rel_max"] <= x_tol
