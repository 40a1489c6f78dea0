"""Host-side rules of the multi-GPU handle (meso_b200/csrc/gang.cu), checked without a GPU: the processor grid the gang picks
for N bricks (LAMMPS' smallest-surface rule, src/comm.cpp:201-287 -- the grid bench.py and the oracle's worlds use) and the
brick every atom is dealt to at upload (the uniform split of Domain::set_local_box, the same `lo <= x < hi` test
tests/mgpu_check.py applies when one process per GPU uploads its own atoms)."""
import ctypes as C
import itertools

import numpy as np

from bench import procgrid_for
from meso_b200 import lib


def layout(ndev, lo, hi, periodic, x=None):
    L = lib.load()
    x = np.zeros((0, 3)) if x is None else np.ascontiguousarray(x, np.float64)
    grid = (C.c_int * 3)()
    owner = np.empty(len(x), np.int32)
    rc = L.meso_gang_layout(ndev, (C.c_double * 3)(*lo), (C.c_double * 3)(*hi), (C.c_int * 3)(*periodic), grid, len(x),
                            x.ctypes.data_as(C.c_void_p), owner.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return tuple(grid), owner


def test_grid_is_lammps_smallest_surface_rule():
    for n in (1, 2, 3, 4, 6, 8, 12, 16, 27):
        assert layout(n, (0, 0, 0), (200, 200, 200), (1, 1, 1))[0] == procgrid_for(n), n
    # a slab: brute force over the factorisations, first minimum in (px, py) order wins
    prd = (120.0, 60.0, 20.0)
    for n in (2, 4, 6, 8):
        best = None
        for px, py in itertools.product(range(1, n + 1), repeat=2):
            if n % px or (n // px) % py:
                continue
            pz = n // px // py
            s = prd[0] * prd[1] / (px * py) + prd[0] * prd[2] / (px * pz) + prd[1] * prd[2] / (py * pz)
            if best is None or s < best[0]:
                best = (s, (px, py, pz))
        assert layout(n, (0, 0, 0), prd, (1, 1, 0))[0] == best[1], n


def test_atoms_are_dealt_by_the_uniform_split():
    rng = np.random.default_rng(7)
    dims = np.array([12.0, 9.0, 14.0])
    x = rng.random((20000, 3)) * dims
    for n in (2, 4, 8):
        grid, owner = layout(n, (0, 0, 0), dims, (1, 1, 1), x)
        seen = np.zeros(len(x), int)
        for r in range(n):
            loc = (r // (grid[1] * grid[2]), (r // grid[2]) % grid[1], r % grid[2])
            lo = np.array([dims[d] * (loc[d] * (1.0 / grid[d])) for d in range(3)])
            hi = np.array([dims[d] * ((loc[d] + 1) * (1.0 / grid[d])) if loc[d] < grid[d] - 1 else dims[d] for d in range(3)])
            mine = np.all((x >= lo) & (x < hi), axis=1)
            assert np.array_equal(owner == r, mine), (n, r)
            seen += mine
        assert np.all(seen == 1)


def test_faces_images_and_walls():
    # points exactly on brick faces belong to the upper brick; periodic images wrap; beyond a wall the outer brick keeps the atom
    x = np.array([[6.0, 1.0, 1.0], [12.0, 1.0, 1.0], [-0.25, 1.0, 1.0], [11.999999999, 1.0, 1.0], [1.0, 1.0, -0.5], [1.0, 1.0, 12.5],
                  [0.0, 0.0, 0.0], [5.999999999999, 1.0, 1.0]])
    grid, owner = layout(2, (0, 0, 0), (12, 12, 12), (1, 1, 1), x)
    assert grid == (1, 1, 2)
    grid, owner = layout(2, (0, 0, 0), (12, 6, 6), (1, 1, 1), x[:, [0, 1, 2]])
    assert grid == (2, 1, 1) and owner.tolist()[:4] == [1, 0, 1, 1] and owner[6] == 0 and owner[7] == 0
    grid, owner = layout(2, (0, 0, 0), (6, 6, 12), (1, 1, 0), np.array([[1.0, 1.0, -0.5], [1.0, 1.0, 12.5], [1.0, 1.0, 6.0], [7.0, 1.0, 5.9]]))
    assert grid == (1, 1, 2) and owner.tolist() == [0, 1, 1, 0]
