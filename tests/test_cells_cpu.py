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


def test_host_threshold_logic_matches_oracle():
    """The product's host-side pieces (no GPU needed): default size threshold (grouped KDE), default marker threshold
    from a histogram, half-even rounding, and the vectorised per-cell decisions."""
    from deepliif_b200 import postprocessing as P
    rng = np.random.default_rng(0)
    for t in range(120):
        n = int(rng.integers(1, 5000))
        m = [rng.integers(0, 256, n), rng.integers(0, 3, n), np.clip(rng.normal(100, 40, n), 0, 255).astype(int),
             np.concatenate([np.zeros(n, int), rng.integers(250, 256, max(1, n // 500))])][t % 4].astype(np.uint8)
        assert C.calculate_default_marker_threshold(m) == P.calculate_default_marker_threshold_from_hist(np.bincount(m, minlength=256))
    for t in range(40):
        n = int(rng.integers(0, 1500))
        sizes = rng.integers(5, 3000, n) if t % 2 else (rng.gamma(2.0, 60, n).astype(np.int64) + 5)
        for res in ("40x", "20x", "10x"):
            assert C.calculate_default_size_threshold(sizes, res) == P.calculate_default_size_threshold(sizes, res)
    num, den = rng.integers(0, 10**12, 5000), rng.integers(1, 10**6, 5000)
    den[:100] = 2; num[:100] = rng.integers(0, 1000, 100)                 # exact ties
    assert P._round_div_half_even(num, den).tolist() == [int(round(int(a) / int(b))) for a, b in zip(num, den)]
    for t in range(30):
        n = int(rng.integers(0, 400))
        tab = P.CellTable()
        tab.count = rng.integers(1, 500, n); tab.positive = rng.random(n) < 0.5; tab.marker = rng.integers(0, 400, n)
        tab.x0 = tab.y0 = tab.sum_x = tab.sum_y = np.zeros(n, np.int64)
        kept = np.sort(rng.choice(2 * n + 1, n, replace=False)) if n else np.zeros(0, np.int64)
        kw = [dict(), dict(marker_thresh=200), dict(size_thresh_upper=300), dict(od_thresh_lower=50, od_thresh_upper=350),
              dict(od_thresh_upper=100, marker_thresh=20)][t % 5]
        args = dict(size_thresh=int(rng.integers(0, 100)), marker_thresh=None, size_thresh_upper=None, od_thresh_lower=None,
                    od_thresh_upper=None); args.update(kw)
        cls, counts = P._classes(tab, kept, 2 * n + 1, **args)
        lab = np.zeros((1, 1), np.int64)                                   # oracle loop on a dummy image: decisions only
        want = np.zeros(2 * n + 2, np.uint8); pos = neg = 0
        for cell, ci in zip(tab.as_tuples(), kept.tolist()):
            if cell[0] > args["size_thresh"] and (args["size_thresh_upper"] is None or cell[0] < args["size_thresh_upper"]):
                is_pos = bool(cell[1])
                if args["marker_thresh"] is not None and cell[2] > args["marker_thresh"]:
                    is_pos = True
                if args["od_thresh_lower"] is not None and cell[2] < args["od_thresh_lower"]:
                    is_pos = False
                elif args["od_thresh_upper"] is not None and cell[2] > args["od_thresh_upper"]:
                    is_pos = False
                want[ci] = 2 if is_pos else 1; pos += is_pos; neg += not is_pos
        assert cls.tolist() == want[:cls.shape[0]].tolist() and counts == {"num_total": pos + neg, "num_pos": pos, "num_neg": neg}
