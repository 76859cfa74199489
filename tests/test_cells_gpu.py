"""Cell post-processing on the GPU (csrc/cells.cu through the C ABI) against the oracle restatement and the reference's
golden outputs.  Everything here is integer work: the bar is bit-exact."""
import numpy as np
import pytest
import torch

from oracle import cells as C
from test_cells_cpu import CELL_CASES, case_inputs, check_against_golden

pytestmark = pytest.mark.gpu


def _gpu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("ci", range(len(CELL_CASES)))
def test_compute_final_results_matches_reference_golden(ci):
    from deepliif_b200 import postprocessing as P
    orig, seg, marker, res, kw = case_inputs(ci)
    overlay, refined, scoring, mask, cells = P.compute_final_results(orig, seg, marker, res, return_mask=True, **kw)
    check_against_golden(ci, overlay, refined, scoring, cells)
    st = {}
    C.compute_final_results(orig, seg, marker, res, stages=st, **kw)
    assert np.array_equal(mask, st["mask"])


def _patterns():
    rng = np.random.default_rng(7)
    out = {}
    for name, (H, W) in {"odd": (37, 45), "row": (1, 130), "col": (131, 1), "one": (1, 1), "wide": (9, 1000)}.items():
        out["rand_" + name] = rng.choice(np.array([50, 150, 200], np.uint8), size=(H, W), p=[0.5, 0.25, 0.25])
    cb = np.full((64, 96), 50, np.uint8); cb[::2, ::2] = 200; cb[1::2, 1::2] = 150      # diagonal-only contacts
    out["checker"] = cb
    out["all_unknown"] = np.full((40, 70), 50, np.uint8)
    out["all_pos"] = np.full((40, 70), 200, np.uint8)
    sp = np.full((101, 103), 200, np.uint8)                                               # spiral corridor of UNKNOWN
    y0, x0, y1, x1 = 0, 0, 100, 102
    while y1 - y0 > 3 and x1 - x0 > 3:
        sp[y0, x0:x1 + 1] = 50; sp[y0:y1 + 1, x1] = 50; sp[y1, x0 + 2:x1 + 1] = 50; sp[y0 + 2:y1 + 1, x0 + 2] = 50
        y0 += 2; x0 += 2; y1 -= 2; x1 -= 2
        sp[y0, x0] = 50
        y0 += 2; x0 += 2; y1 -= 2; x1 -= 2
    out["spiral"] = sp
    comb = np.full((90, 300), 50, np.uint8); comb[5:85:2, 3:297] = 200; comb[5:85, 3] = 150   # long thin teeth, one cell
    out["comb"] = comb
    rings = np.full((120, 120), 50, np.uint8)
    for r in range(4, 58, 4):
        rings[60 - r:60 + r + 1, 60 - r] = 200; rings[60 - r:60 + r + 1, 60 + r] = 200
        rings[60 - r, 60 - r:60 + r + 1] = 150; rings[60 + r, 60 - r:60 + r + 1] = 150
    out["rings"] = rings                                                                      # nested enclosed regions
    dense = rng.choice(np.array([50, 150, 200], np.uint8), size=(257, 513), p=[0.6, 0.2, 0.2])
    out["dense"] = dense
    return out


PATTERNS = _patterns()


@pytest.mark.parametrize("name", sorted(PATTERNS))
def test_mark_background_and_labels_match_oracle(name):
    from deepliif_b200 import postprocessing as P
    m0 = PATTERNS[name]
    H, W = m0.shape
    want_bg = C.mark_background(m0)
    got = P.mark_background_gpu(_gpu(m0).clone())
    assert np.array_equal(got.cpu().numpy(), want_bg)
    rng = np.random.default_rng(H * 1000 + W)
    marker = rng.integers(0, 65535, size=(H, W)).astype(np.uint16)
    for use_avg in (False, True):
        lab_o, t_o = C.components(want_bg, marker, use_avg)
        labels, roots, n = P.label_cells_gpu(got)
        assert n == t_o.shape[0]
        assert np.array_equal(labels.cpu().numpy(), lab_o.astype(np.int32) - 1)
        t = P.cell_stats_gpu(got, _gpu(marker.view(np.int16)).view(torch.uint16), labels.reshape(-1), roots, use_avg).cpu().numpy()
        assert np.array_equal(t, t_o)
        if n:
            assert np.array_equal(roots.cpu().numpy(), t_o[:, 5] * W + t_o[:, 4])


@pytest.mark.parametrize("name", ["rand_odd", "checker", "rings", "dense", "comb"])
def test_classify_enlarge_final_match_oracle(name):
    from deepliif_b200 import postprocessing as P
    m0 = PATTERNS[name]
    H, W = m0.shape
    bg = C.mark_background(m0)
    mask_c, cells, lab, kept = C.compute_cell_mapping(bg, None, 0, None)
    rng = np.random.default_rng(5)
    n = int(lab.max())
    cls = rng.integers(0, 3, size=max(n, 1)).astype(np.uint8)
    # oracle classification driven by the same per-cell classes: express them as marker-threshold-free cells
    sel = [i for i in range(len(cells)) if cls[kept[i]] > 0]
    cells_o = [(cells[i][0], cls[kept[i]] == 2) + tuple(cells[i][2:]) for i in sel]
    want, _ = C.create_cell_classification(mask_c, cells_o, lab, kept[sel], size_thresh=-1)
    labels, roots, n_g = P.label_cells_gpu(_gpu(bg))
    assert n_g == n
    got = P.classify_gpu(labels.reshape(-1), roots, _gpu(cls), H, W)
    assert np.array_equal(got.cpu().numpy(), want)
    e1 = P.enlarge_cell_boundaries_gpu(got)
    assert np.array_equal(e1.cpu().numpy(), C.enlarge_cell_boundaries(want))
    e2 = P.enlarge_cell_boundaries_gpu(e1)
    orig = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    ov, rf = P.create_final_images_gpu(_gpu(orig), e2)
    ov_o, rf_o = C.create_final_images(orig, C.enlarge_cell_boundaries(C.enlarge_cell_boundaries(want)))
    assert np.array_equal(ov.cpu().numpy(), ov_o) and np.array_equal(rf.cpu().numpy(), rf_o)


def test_marker_plane_and_posneg_mask_bit_exact():
    from deepliif_b200 import postprocessing as P
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=(123, 77, 3), dtype=np.uint8)
    img[:5] = 0
    plane, hist = P.marker_plane_gpu(_gpu(img), use_od=False, want_hist=True)
    gray = img.max(axis=-1)
    assert np.array_equal(plane.view(torch.int16).cpu().numpy().view(np.uint16), gray.astype(np.uint16))
    want_hist = np.bincount(gray.ravel(), minlength=256); want_hist[0] = 0
    assert np.array_equal(hist.cpu().numpy(), want_hist)
    od, _ = P.marker_plane_gpu(_gpu(img), use_od=True)
    assert np.array_equal(od.view(torch.int16).cpu().numpy().view(np.uint16), C.create_od_image(img))
    # every (r, g, b) grey level pair the LUT sum can tie on: the full 256^2 plane with b = r
    rr, gg = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8))
    full = np.stack([rr, gg, rr], axis=-1)
    od, _ = P.marker_plane_gpu(_gpu(full), use_od=True)
    assert np.array_equal(od.view(torch.int16).cpu().numpy().view(np.uint16), C.create_od_image(full))
    for thresh in (0, 90, 120, 300):
        m = P.create_posneg_mask_gpu(_gpu(img), thresh)
        assert np.array_equal(m.cpu().numpy(), C.create_posneg_mask(img, thresh))


def test_full_size_region_properties_and_oracle():
    """A 3000 x 4000 region (12 MP, the order of one WSI region): whole pipeline against the oracle."""
    from deepliif_b200 import postprocessing as P
    orig, seg, marker = C.synth_case(3000, 4000, 21)
    ov, rf, sc, mask, cells = P.compute_final_results(orig, seg, marker, "40x", marker_thresh="default",
                                                      large_noise_thresh="default", return_mask=True)
    st = {}
    ov_o, rf_o, sc_o = C.compute_final_results(orig, seg, marker, "40x", marker_thresh="default",
                                               large_noise_thresh="default", stages=st)
    assert sc == sc_o and len(cells) == len(st["cells"]) and cells == [tuple(c) for c in st["cells"]]
    assert np.array_equal(mask, st["mask"]) and np.array_equal(ov, ov_o) and np.array_equal(rf, rf_o)
    assert sc["num_total"] > 1000


def test_dims_rejected_loudly():
    from deepliif_b200 import _lib
    lib = _lib.load()
    assert lib.dlb_cells_enlarge(None, None, 65536, 65536, None) != 0
    assert b"2^31" in lib.dlb_last_error()
