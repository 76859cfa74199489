"""Cell post-processing oracle (oracle/cells.py) against the reference's own numba results (tests/golden/cells.npz,
written by oracle/gen_golden.py from deepliif/postprocessing.py:193-308, 923-1071, 1223-1304)."""
import json
import os

import numpy as np
import pytest

from oracle import cells as C
from oracle.gen_golden import CELL_CASES

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "cells.npz"))


def case_inputs(ci):
    H, W, seed, res, kw = CELL_CASES[ci]
    kw = dict(kw)
    orig, seg, marker = C.synth_case(H, W, seed)
    if kw.pop("no_marker", False):
        marker = None
    return orig, seg, marker, res, kw


def check_against_golden(ci, overlay, refined, scoring, cells=None, bg=None):
    g = lambda k: GOLD[f"c{ci}_{k}"]
    assert json.dumps(scoring, sort_keys=True) == bytes(g("scoring")).decode()
    if bg is not None:
        assert np.array_equal(bg, g("bg"))
    if cells is not None:
        got = np.array([[c[0], int(c[1]), int(c[2]), c[3], c[4], c[5], c[6]] for c in cells], dtype=np.int64).reshape(-1, 7)
        assert np.array_equal(got, g("cells"))
    if f"c{ci}_overlay" in GOLD:
        assert np.array_equal(overlay, g("overlay")) and np.array_equal(refined, g("refined"))
    else:
        W = refined.shape[1]
        assert np.array_equal(refined[::3, ::5], g("refined_sub"))
        sums = [int(overlay.astype(np.int64).sum()), int(refined.astype(np.int64).sum()),
                int((refined.astype(np.int64) * (np.arange(W)[None, :, None] + 1)).sum())]
        assert sums == g("sums").tolist()


@pytest.mark.parametrize("ci", range(len(CELL_CASES)))
def test_cells_oracle_matches_reference_golden(ci):
    orig, seg, marker, res, kw = case_inputs(ci)
    st = {}
    overlay, refined, scoring = C.compute_final_results(orig, seg, marker, res, stages=st, **kw)
    bg = C.mark_background(C.create_posneg_mask(seg, kw.get("seg_thresh", 120)))
    check_against_golden(ci, overlay, refined, scoring, st["cells"], bg)
    d = GOLD[f"c{ci}_defaults"].tolist()
    assert st["defaults"].get("size_thresh", -1) == d[0] and st["defaults"].get("marker_thresh", -1) == d[1]


def test_mark_background_fixed_point_properties():
    """Size-independent properties: idempotent, never touches classified pixels, leaves only enclosed UNKNOWN."""
    _, seg, _ = C.synth_case(300, 300, 99)
    m0 = C.create_posneg_mask(seg, 120)
    m1 = C.mark_background(m0)
    assert np.array_equal(C.mark_background(m1), m1)
    assert np.array_equal(m1[m0 != C.LABEL_UNKNOWN], m0[m0 != C.LABEL_UNKNOWN])
    assert not (m1[0] == 50).any() and not (m1[-1] == 50).any() and not (m1[:, 0] == 50).any() and not (m1[:, -1] == 50).any()
    left = m1 == C.LABEL_UNKNOWN                          # what is left has no BACKGROUND 4-neighbour
    bgm = m1 == 0
    nb = np.zeros_like(bgm)
    nb[1:] |= bgm[:-1]; nb[:-1] |= bgm[1:]; nb[:, 1:] |= bgm[:, :-1]; nb[:, :-1] |= bgm[:, 1:]
    assert not (left & nb).any()
