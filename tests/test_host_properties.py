"""Property tests of the host-side geometry and partitioning logic (hypothesis; no GPU)."""
import math

import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

import ceg_b200 as ceg
from ceg_b200 import sharding


cells = st.tuples(st.floats(4.0, 40.0), st.floats(4.0, 40.0), st.floats(4.0, 40.0),
                  st.floats(62.0, 118.0), st.floats(62.0, 118.0), st.floats(62.0, 118.0))


def _mat(c):
    m = ceg.mat_from_parameters(c[:3], c[3:])
    return m if np.isfinite(m).all() and abs(np.linalg.det(m)) > 1.0 else None


@given(st.integers(1, 5000), st.integers(1, 64))
def test_split_range_partitions(n, parts):
    r = sharding.split_range(n, parts)
    assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    sizes = [b - a for a, b in r]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@given(st.integers(1, 40), st.integers(1, 9), st.integers(1, 9), st.integers(1, 8), st.booleans())
def test_slab_file_runs_tile_the_payload(nx, ny, nz, world, coulomb):
    """The 8 runs of every rank's x-slab are disjoint and together cover the .grid payload exactly."""
    header = 136 if coulomb else 128
    covered = np.zeros(header + 8 * nx * ny * nz * 4, dtype=np.int32)
    for (a, b) in sharding.split_range(nx, world):
        if a == b:
            continue
        for off, length in sharding.slab_file_runs((nx, ny, nz), a, b, header):
            covered[off:off + length] += 1
    assert (covered[:header] == 0).all() and (covered[header:] == 1).all()


@settings(max_examples=60, deadline=None)
@given(cells, st.floats(0.08, 0.6))
def test_grid_coordinates_setup(c, spacing):
    """coordinates.jl:32-41: odd dims, delta = size / dims, and the box contains the whole cell."""
    m = _mat(c)
    if m is None:
        return
    cs = ceg.GridCoordinatesSetup.from_cell(m, spacing)
    dims = np.asarray(cs.dims)
    assert (dims % 2 == 1).all() and (dims >= 1).all()
    assert np.allclose(np.asarray(cs.delta) * dims, cs.size, rtol=1e-15, atol=0)
    raw = np.floor(np.asarray(cs.size) / spacing).astype(int)
    assert np.array_equal(dims, raw + (raw % 2 == 0))                    # floor, then bumped to the next odd number
    corners = np.array([[i, j, k] for i in (0, 1) for j in (0, 1) for k in (0, 1)], dtype=float) @ m.T
    lo, hi = np.asarray(cs.shift), np.asarray(cs.shift) + np.asarray(cs.size)
    assert (corners >= lo - 1e-9).all() and (corners <= hi + 1e-9).all()


@settings(max_examples=60, deadline=None)
@given(cells, st.tuples(st.floats(0.01, 0.99), st.floats(0.01, 0.99), st.floats(0.01, 0.99)),
       st.tuples(st.integers(-2, 2), st.integers(-2, 2), st.integers(-2, 2)))
def test_offsetpoint_is_periodic_and_invertible(c, frac, shift):
    """coordinates.jl:58-70: the grid coordinate of a point does not depend on the periodic image and maps back."""
    m = _mat(c)
    if m is None:
        return
    cs = ceg.GridCoordinatesSetup.from_cell(m, 0.3)
    p = m @ np.array(frac)
    s = ceg.offsetpoint(p, cs)
    assert (s >= 1.0 - 1e-9).all() and (s <= np.asarray(cs.dims) + 1.0 + 1e-9).all()
    assert np.allclose(ceg.inverse_offsetpoint(s, cs), p, atol=1e-8)
    assert np.allclose(ceg.offsetpoint(p + m @ np.array(shift, dtype=float), cs), s, atol=1e-7)


@settings(max_examples=80, deadline=None)
@given(cells, st.floats(6.0, 14.0))
def test_find_supercell_is_minimal(c, cutoff):
    """utils.jl:10-67: every perpendicular width of the supercell reaches 2 x cutoff, with no axis over-replicated."""
    m = _mat(c)
    if m is None:
        return
    sc = ceg.find_supercell(m, cutoff)
    a, b, cc = m[:, 0], m[:, 1], m[:, 2]
    vol = abs(np.linalg.det(m))
    widths = np.array([vol / np.linalg.norm(np.cross(b, cc)), vol / np.linalg.norm(np.cross(cc, a)),
                       vol / np.linalg.norm(np.cross(a, b))])
    for d in range(3):
        assert sc[d] >= 1
        assert sc[d] * widths[d] >= 2 * cutoff * (1 - 1e-9)
        assert sc[d] == 1 or (sc[d] - 1) * widths[d] < 2 * cutoff * (1 + 1e-9)


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 4), st.integers(1, 4), st.integers(1, 4), st.booleans(), st.integers(0, 2 ** 31 - 1))
def test_grid_file_roundtrip_random(ax, ay, az, coulomb, seed):
    """grids.jl:61-116: write_grid -> parse_grid returns the header and GRID_TO_KELVIN * payload (as Float32)."""
    rng = np.random.default_rng(seed)
    m = ceg.mat_from_parameters((3.0 + ax, 3.5 + ay, 4.0 + az), (80.0 + 5 * ax, 95.0, 100.0 - 3 * az))
    cs = ceg.GridCoordinatesSetup.from_cell(m, 0.9)
    nx, ny, nz = cs.nodes
    g = rng.standard_normal((8, nx, ny, nz)).astype(np.float32)
    import os
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "x.grid")
        ceg.write_grid(path, cs, (1, 2, 3), g, ewald_precision=1e-6 if coulomb else None)
        assert os.path.getsize(path) == (136 if coulomb else 128) + g.nbytes + 72
        eg = ceg.parse_grid(path, coulomb)
    assert eg.num_unitcell == (1, 2, 3) and np.array_equal(eg.csetup.dims, cs.dims)
    assert eg.ewald_precision == (1e-6 if coulomb else math.inf)
    assert np.array_equal(eg.grid, (g.astype(np.float64) * ceg.GRID_TO_KELVIN).astype(np.float32))
    assert np.allclose(eg.csetup.cell.mat, m, rtol=0, atol=0)
