"""Host logic of the zoom levels (genetic_b200/zoom.py: selection, window placement, flag transfer; no field data, runs
without a GPU) against the oracle's independent restatement and the reference's golden grid geometry."""
import numpy as np
import pytest

import genetic_b200 as gb
from genetic_b200 import multilevel as GM, zoom as Z
from oracle import multilevel as ML


def _pair(box, n):
    ctx = GM.MultiLevelGrid()
    ctx.addLevel(gb.Grid(box, n))
    return ctx, [ML.Grid(box, n)]


def _same(ctx, og):
    for g, o in zip(ctx.grids, og):
        assert g.size == o.n and float(g.cellSize) == o.cell
        assert tuple(float(v) for v in g.offsetLower) == o.offset
        assert np.array_equal(g.getFlaggedCells(), o.flags)


def test_two_zooms_golden_geometry_01d():
    """test_01d: basegrid 50 64; centre 25 25 25; select_sphere 5; zoomgrid 3 64; select_sphere 3; zoomgrid 2 64.
    reference_grid/grid-info-{1,2}.txt: offsets 16.4062 / 20.5729, sizes 16.6667 / 8.33333."""
    ctx, og = _pair(50.0, 64)
    Z.selectSphere(ctx, (25, 25, 25), 5)
    ML.select_sphere(og, (25, 25, 25), 5)
    Z.initZoomGrid(ctx, 3, 64)
    ML.add_zoom_level(og, 3, 64)
    Z.selectSphere(ctx, (25, 25, 25), 3)
    ML.select_sphere(og, (25, 25, 25), 3)
    Z.initZoomGrid(ctx, 2, 64)
    ML.add_zoom_level(og, 2, 64)
    _same(ctx, og)
    assert "%.6g" % float(ctx.grids[1].offsetLower[0]) == "16.4062" and "%.6g" % float(ctx.grids[1].thisGridSize) == "16.6667"
    assert "%.6g" % float(ctx.grids[2].offsetLower[0]) == "20.5729" and "%.6g" % float(ctx.grids[2].thisGridSize) == "8.33333"


@pytest.mark.parametrize("box,n,centre,sel,zoom", [
    (64.0, 64, (25, 32, 24), ("cube", 16), (2, 64)),      # test_02m
    (50.0, 32, (1, 10, 25), ("sphere", 5.0), (4, 32)),    # test_01e: the window wraps around the box
    (16.0, 16, (8, 8, 8), ("cube", 4), (4, 16)),          # test_23
    (50.0, 64, (49.5, 0.2, 25), ("sphere", 4.0), (4, 64)),
])
def test_selection_and_window_placement_match_oracle(box, n, centre, sel, zoom):
    ctx, og = _pair(box, n)
    if sel[0] == "cube":
        Z.selectCube(ctx, centre, sel[1]); ML.select_cube(og, centre, sel[1])
    else:
        Z.selectSphere(ctx, centre, sel[1]); ML.select_sphere(og, centre, sel[1])
    assert Z.zoomCorner(ctx.grids[0], zoom[0]) == ML.zoom_corner(og[0], zoom[0])
    Z.initZoomGrid(ctx, *zoom)
    ML.add_zoom_level(og, *zoom)
    _same(ctx, og)
    Z.selectNearest(ctx, centre)
    ML.select_nearest(og, centre)
    _same(ctx, og)


def test_errors():
    ctx, _ = _pair(50.0, 32)
    with pytest.raises(gb.GxError):
        Z.initZoomGrid(ctx, 2, 32)                       # nothing flagged
    Z.selectSphere(ctx, (25, 25, 25), 20)
    with pytest.raises(gb.GxError):
        Z.initZoomGrid(ctx, 4, 32)                       # selection larger than the sub-box
