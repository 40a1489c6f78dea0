"""CPU tests of the oracle's whole-step restatement: pinned against stock LAMMPS (golden fixture made by
tests/golden/make_lammps_golden.py with the reference tree's own pair_style dpd), and checked for the
size-independent properties of the domain (momentum conservation, decomposition independence of the
neighbor SETS, periodic images, tile-transposed layout)."""
import os

import numpy as np
import pytest

import oracle
from meso_b200 import workload

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "stock_dpd_conservative_L8.npz"))


def world(L, precision=0, procgrid=(1, 1, 1), gamma_sigma=True, x=None, v=None, **kw):
    x = workload.dpd_fluid(L) if x is None else x
    v = workload.maxwell_velocities(len(x)) if v is None else v
    g, s = (4.5, 3.0) if gamma_sigma else (0.0, 0.0)
    dims = (L, L, L) if np.isscalar(L) else L
    w = oracle.World((0, 0, 0), dims, procgrid=procgrid, precision=precision, coeff=oracle.default_coeff(1, 15.0, g, s), **kw)
    w.set_atoms(x, v)
    return w


def by_tag(w, key="f"):
    n = sum(w.counts(r)["nlocal"] for r in range(w.nranks))
    out = np.zeros((n, 3))
    for r in range(w.nranks):
        a = w.atoms(r)
        out[a["tag"][:a["nlocal"]] - 1] = a[key][:a["nlocal"]]
    return out


@pytest.mark.parametrize("precision,tol", [(1, 1e-5), (0, 1e-5)])
def test_conservative_forces_match_stock_lammps(precision, tol):
    """gamma = sigma = 0 against stock pair_style dpd.  Stock LAMMPS works on fp64 coordinates, the MESO
    styles on fl32(x - centre), so even the fp64 style is limited by the packing (~1e-6)."""
    w = world(int(GOLD["L"]), precision, gamma_sigma=False)
    w.setup(eflag=1, vflag=1)
    f = by_tag(w)
    ref = GOLD["f"]
    mag = np.linalg.norm(ref, axis=1)
    err = (np.linalg.norm(f - ref, axis=1) / np.maximum(mag, mag.mean())).max()
    assert err <= tol, err
    vir, e = w.virial()
    n = len(ref)
    assert abs(e.sum() / n - float(GOLD["pe"])) < 1e-5 * float(GOLD["pe"])
    # P = (N kT + sum_i tr(W_i)) / (3V); thermo `press` at run 0 with the T = 1 velocities is not stored, so
    # check the virial part only: P_virial = sum(vxx+vyy+vzz) / (3 V)
    L = float(GOLD["L"])
    p_vir = vir[:, :3].sum() / (3 * L ** 3)
    assert abs(p_vir - float(GOLD["press"])) < 1e-4 * float(GOLD["press"])   # golden run had v = 0: kinetic part is 0


def test_momentum_conservation_and_symmetry():
    for precision in (0, 1):
        w = world(8, precision)
        w.setup()
        f = by_tag(w)
        # pairs across a periodic face see fl32-rounded image coordinates: antisymmetry holds to ~1e-7 per pair
        assert np.abs(f.sum(0)).max() < 5e-3
        cnt, rows = w.neighbors()
        a = w.atoms()
        tags = a["tag"]
        sets = [set(tags[rows[i, :cnt[i]]]) for i in range(a["nlocal"])]
        loc = {int(t): i for i, t in enumerate(tags[:a["nlocal"]])}
        for i in range(0, a["nlocal"], 7):
            for t in sets[i]:
                assert int(tags[i]) in sets[loc[int(t)]]


def test_neighbor_order_core_then_reversed_skin():
    w = world(6, 1)
    w.setup()
    cnt, rows = w.neighbors()
    c4, _ = w.packed()
    for i in range(0, len(cnt), 11):
        j = rows[i, :cnt[i]]
        d = c4[i, :3] - c4[j, :3]
        r2 = d[:, 2] * d[:, 2] + (d[:, 1] * d[:, 1] + d[:, 0] * d[:, 0])
        core = r2 <= 1.0
        ncore = int(core.sum())
        assert core[:ncore].all() and not core[ncore:].any()
        assert (r2 <= np.float32(1.69) * (1 + 1e-6)).all()


def test_transposed_layout_roundtrip():
    w = world(6, 0)
    w.setup()
    cnt, rows = w.neighbors()
    t = w.neighbors_transposed()
    n_col = w.counts()["n_col"]
    for i in (0, 1, 31, 32, 33, len(cnt) - 1):
        for k in range(cnt[i]):
            assert t[((i & ~31) + (k & 31)) * n_col + (k >> 5) * 32 + (i & 31)] == rows[i, k]


@pytest.mark.parametrize("procgrid", [(2, 1, 1), (1, 2, 2), (2, 2, 2), (3, 1, 1)])
def test_decomposition_independent_neighbor_sets_and_forces(procgrid):
    """bit-exactness of the ordered list holds per decomposition (the packing centre moves); the neighbor SETS
    by tag and the fp64 forces must not depend on the processor grid."""
    L = 12
    x, v = workload.dpd_fluid(L), workload.maxwell_velocities(4 * L ** 3)
    w1 = world(L, 1, x=x, v=v)
    wn = world(L, 1, procgrid=procgrid, x=x, v=v)
    w1.setup(); wn.setup()

    def tagsets(w):
        out = {}
        for r in range(w.nranks):
            a = w.atoms(r)
            cnt, rows = w.neighbors(r)
            for i in range(a["nlocal"]):
                out[int(a["tag"][i])] = frozenset(a["tag"][rows[i, :cnt[i]]].tolist())
        return out

    s1, sn = tagsets(w1), tagsets(wn)
    diff = [t for t in s1 if s1[t] != sn[t]]
    # membership is decided on fl32(x - centre): a pair within 1 ulp of r_n may flip with the centre
    assert len(diff) <= 2, len(diff)
    f1, fn = by_tag(w1), by_tag(wn)
    mag = np.linalg.norm(f1, axis=1)
    assert (np.linalg.norm(f1 - fn, axis=1) / np.maximum(mag, mag.mean())).max() < 5e-5
    assert sum(wn.counts(r)["nlocal"] for r in range(wn.nranks)) == 4 * L ** 3
    w1.run(7); wn.run(7)        # crosses a rebuild with migration
    assert sum(wn.counts(r)["nlocal"] for r in range(wn.nranks)) == 4 * L ** 3
    assert abs(w1.temperature() - wn.temperature()) < 0.05


def test_ghosts_are_periodic_images():
    L = 7
    w = world(L, 0)
    w.setup()
    a = w.atoms()
    nl = a["nlocal"]
    pos = {int(t): a["x"][i] for i, t in enumerate(a["tag"][:nl])}
    for g in range(nl, nl + a["nghost"], 5):
        d = a["x"][g] - pos[int(a["tag"][g])]
        k = np.round(d / L)
        assert np.all(np.isin(k, (-1, 0, 1))) and np.array_equal(a["x"][g], pos[int(a["tag"][g])] + k * L)
        assert np.any(d != 0)
    c = w.counts()
    assert c["n_bulk"] + c["n_border"] == nl
    # every atom within the ghost cutoff of a face is a border atom and comes after all bulk atoms
    xl = a["x"][:nl]
    near = ((xl <= 1.3) | (xl >= L - 1.3)).any(1)
    assert not near[:c["n_bulk"]].any() and near[c["n_bulk"]:].all()


def test_nve_energy_drift_conservative():
    """velocity-Verlet with gamma = sigma = 0 conserves KE + PE"""
    x = workload.dpd_fluid(6, seed=11)
    w = world(6, 1, gamma_sigma=False, x=x, v=workload.maxwell_velocities(len(x), 0.1))

    def etot():
        w.force_clear(); w.pair_compute(1, 1)
        _, e = w.virial()
        a = w.atoms()
        return e.sum() + 0.5 * (a["v"][:a["nlocal"]] ** 2).sum()

    w.setup(1, 1)
    e0 = etot()
    w.run(40, 1, 1)
    e1 = etot()
    assert abs(e1 - e0) < 2e-3 * abs(e0), (e0, e1)


def test_harmonic_bonds_and_exclusion_filter_against_direct_evaluation():
    """the oracle's gpu_bond_harmonic / gpu_filter_exclusion restatement against a direct numpy
    evaluation by tag (minimum image), on bead-spring chains in solvent."""
    from meso_b200 import workload
    L, k, r0 = 6, 50.0, 0.5
    x, typ, tag, nb, bt, ba = workload.polymer_melt(L, chain_len=6, seed=3)
    n = len(x)
    coeff = np.array([[1, 1, 1, 1, 25, 4.5, 3.0]] * 4, dtype=float)

    def world(lj12):
        w = oracle.World((0, 0, 0), (L, L, L), ntypes=2, mass=[0, 1, 1], coeff=coeff, precision=1)
        w.set_atoms(x, np.zeros_like(x), tag=tag, type=typ)
        w.set_bonds(nb, bt, ba, tag=tag, k=[0, k], r0=[0, r0], special_lj12=lj12)
        w.rebuild()
        return w

    w = world(1.0)
    w.force_clear(); w.bond_compute(1, 1)
    a = w.atoms()
    pos = np.empty((n + 1, 3)); pos[a["tag"][:n]] = a["x"][:n]
    E, F = 0.0, np.zeros((n + 1, 3))
    for i in range(n):
        for p in range(nb[i]):
            d = pos[ba[i, p]] - pos[tag[i]]
            d -= L * np.round(d / L)
            r = np.linalg.norm(d)
            E += 0.5 * k * (r - r0) ** 2                      # each bond is held by both atoms: e_bond = e/2 per atom
            F[tag[i]] += 2 * k * (r - r0) / r * d
    Fo = np.empty((n + 1, 3)); Fo[a["tag"][:n]] = a["f"]
    assert abs(w.bond_energy() - E) < 1e-5 * E                # fp32-packed coordinates
    assert np.abs(F[1:] - Fo[1:]).max() < 2e-4 and np.abs(a["f"].sum(0)).max() < 1e-9
    c_all, rows_all = w.neighbors()
    c_ex, rows_ex = world(0.0).neighbors()
    assert c_all.sum() - c_ex.sum() == nb.sum()               # every bond partner is within r_n here and leaves the row
    t = a["tag"]
    for i in np.flatnonzero(nb[a["tag"][:n] - 1] > 0)[:50]:
        partners = set(ba[t[i] - 1, :nb[t[i] - 1]].tolist())
        kept = [j for j in rows_all[i, :c_all[i]] if t[j] not in partners]
        assert kept == rows_ex[i, :c_ex[i]].tolist()          # survivors keep their order


@pytest.mark.parametrize("lj12", [1, 0])
def test_polymer_forces_match_stock_lammps_bond_harmonic(lj12):
    """SURVEY.md s8f N1 pin: conservative pair + harmonic bond forces of bead-spring chains against stock LAMMPS
    (atom_style bond, bond_style harmonic of src/MOLECULE, pair_style dpd gamma = 0; fixture made by
    tests/golden/make_lammps_bond_golden.py), with 1-2 pairs kept (special_bonds lj 1 1 1) and excluded (lj 0 1 1).
    The MESO styles work on fl32(x - centre): 1e-5 relative, as for the plain fluid."""
    from meso_b200 import workload
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "stock_polymer_conservative_L6.npz"))
    L = int(g["L"])
    x, typ, tag, nb, bt, ba = workload.polymer_melt(L, chain_len=int(g["chain_len"]), seed=int(g["seed"]))
    coeff = np.array([[1, 1, 1, 1, 25, 0, 0], [1, 1, 1, 1, 40, 0, 0], [1, 1, 1, 1, 40, 0, 0], [1, 1, 1, 1, 25, 0, 0]], dtype=float)
    for precision in (0, 1):
        w = oracle.World((0, 0, 0), (L, L, L), ntypes=2, mass=[0, 1, 1], coeff=coeff, precision=precision)
        w.set_atoms(x, np.zeros_like(x), tag=tag, type=typ)
        w.set_bonds(nb, bt, ba, tag=tag, k=[0, float(g["k"])], r0=[0, float(g["r0"])], special_lj12=float(lj12))
        w.setup(eflag=1, vflag=1)
        f = by_tag(w)
        ref = g["f_lj%d" % lj12]
        mag = np.linalg.norm(ref, axis=1)
        err = (np.linalg.norm(f - ref, axis=1) / np.maximum(mag, mag.mean())).max()
        assert err <= 1e-5, (precision, err)
        assert abs(w.bond_energy() - float(g["ebond_lj%d" % lj12])) < 1e-5 * float(g["ebond_lj%d" % lj12])
        _, e = w.virial()
        assert abs(e.sum() - float(g["evdwl_lj%d" % lj12])) < 1e-5 * float(g["evdwl_lj%d" % lj12])


def test_amphiphilic_channel_is_decomposition_independent():
    """The checker of the multi-GPU polymer runs: bead-spring chains + 1-2 exclusions + walls + body force on 2x2x2 simulated
    ranks (bonds crossing brick faces, partners among the ghosts, the bond table following migrating atoms, a non-periodic
    decomposed dimension) against the same system on one rank.  Conservative forces only (gamma = sigma = 0): the random
    force is keyed on the fp32 velocity BITS, and coordinates are packed relative to each brick's centre, so with the
    thermostat on a last-bit difference re-draws a pair's random number -- trajectories then agree statistically only."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from test_fixes import AMPHI_COEFF
    from meso_b200 import workload
    Lc = 8
    x, typ, tag, nb, bt, ba = workload.amphiphilic_channel(Lc)
    v = workload.maxwell_velocities(len(x), seed=99) * 2.0               # hot: atoms cross brick faces within 17 steps
    coeff = np.zeros((3, 3, 7))
    for (a, b), a0 in AMPHI_COEFF.items():
        coeff[a - 1, b - 1] = coeff[b - 1, a - 1] = [1.0, 1.0, 1.0, 1.0, a0, 0.0, 0.0]
    out = {}
    for grid in ((1, 1, 1), (2, 2, 2)):
        w = oracle.World((0, 0, 0), (Lc, Lc, Lc), periodic=(1, 1, 0), procgrid=grid, ntypes=3, mass=[0, 1, 1, 1],
                         coeff=coeff.reshape(-1, 7), precision=1)
        w.set_atoms(x, v, tag=tag, type=typ)
        w.set_bonds(nb, bt, ba, tag=tag, k=[0.0, 50.0], r0=[0.0, 0.5], special_lj12=0.0)
        w.fix_solid_bound("z"); w.fix_pois(2, 0, 0.2)
        w.setup()
        n0 = [w.counts(r)["nlocal"] for r in range(w.nranks)]
        w.run(17)
        n1 = [w.counts(r)["nlocal"] for r in range(w.nranks)]
        out[grid] = (by_tag(w, "x"), by_tag(w, "v"), n0, n1)
    (x1, v1, _, _), (x8, v8, n0, n1) = out[(1, 1, 1)], out[(2, 2, 2)]
    assert sum(n1) == len(x) and n0 != n1, "no atom migrated: the test would not exercise the bond table's migration"
    # fp32 packing relative to different centres: forces differ by ~1e-6 relative, positions by ~1e-7 after 17 steps
    assert np.abs(x1 - x8).max() < 1e-5 and np.abs(v1 - v8).max() < 1e-3, (np.abs(x1 - x8).max(), np.abs(v1 - v8).max())


@pytest.mark.parametrize("case", ["cube", "brick_of_8", "channel"])
def test_block_segments_and_tile_sizes_of_the_tile_build(case):
    """Design invariants of the neighbor build (meso_b200/csrc/neighbor.cu:k_build_tiles), checked on the oracle's own sorted atoms
    and cells: (1) the local atoms of an aligned 4 x 4 x 4 block of cells are ONE run of consecutive indices per class (bulk /
    border), so a box has about as many segments as (block, class) pairs -- never one per cell; (2) the tile of a block (its
    cells plus one layer, every x-row rounded out to whole 16-byte groups) stays inside the shared-memory budget the host-side
    geometry reserves for this density; (3) the candidates of every atom's 27 stencil cells lie in its block's tile."""
    if case == "cube":
        w, r = world(16), 0
    elif case == "brick_of_8":
        w, r = world(24, procgrid=(2, 2, 2)), 5
    else:
        x = workload.dpd_fluid(16)
        w = oracle.World((0, 0, 0), (16, 16, 16), periodic=(1, 1, 0))
        w.set_atoms(x, workload.maxwell_velocities(len(x)))
        r = 0
    w.setup()
    a = w.atoms(r)
    nl = a["nlocal"]
    m, _, _ = w.bins(r)
    cs, ca = w.cells(r)
    cell_of = np.empty(len(ca), np.int64)
    for c in range(len(cs) - 1):
        cell_of[ca[cs[c]:cs[c + 1]]] = c
    cx, cy, cz = cell_of % m[0], (cell_of // m[0]) % m[1], cell_of // (m[0] * m[1])
    block = (cx >> 2) | ((cy >> 2) << 10) | ((cz >> 2) << 20)
    # (1) segments = maximal runs of equal block id among the locals, in index order
    heads = np.flatnonzero(np.r_[True, block[1:nl] != block[:nl - 1]])
    n_bulk = w.counts(r)["n_bulk"]
    pairs = len(set(block[:n_bulk].tolist())) + len(set(block[n_bulk:nl].tolist()))
    assert len(heads) == pairs, (len(heads), pairs)
    assert len(heads) < 0.2 * len(set(cell_of[:nl].tolist()))             # far fewer segments than occupied cells
    # (2) tile records of every block, rows rounded out to groups of four records
    inner = max(m[0] - 2, 1) * max(m[1] - 2, 1) * max(m[2] - 2, 1)
    per_cell = nl / inner
    cap = (int(216 * per_cell * 1.15) + 4 * 36 + 32 + 63) // 64 * 64       # tile_geometry() of neighbor.cu for 4 x 4 x 4 blocks
    assert cap <= 2560
    worst = 0
    for b in set(block[:nl].tolist()):
        X0, Y0, Z0 = (b & 1023) << 2, ((b >> 10) & 1023) << 2, (b >> 20) << 2
        xlo, xhi = max(X0 - 1, 0), min(X0 + 4, m[0] - 1)
        total = 0
        for z in range(max(Z0 - 1, 0), min(Z0 + 4, m[2] - 1) + 1):
            for y in range(max(Y0 - 1, 0), min(Y0 + 4, m[1] - 1) + 1):
                c0 = xlo + m[0] * (y + m[1] * z)
                g0, g1 = cs[c0] & ~3, (cs[c0 + xhi - xlo + 1] + 3) & ~3
                total += g1 - g0
        worst = max(worst, total)
    assert worst <= cap, (worst, cap)
    # (3) every stencil cell of every local atom lies inside its block's tile (block +- one cell, clipped to the lattice)
    for i in np.random.default_rng(3).choice(nl, size=200, replace=False):
        for c in w.stencil(int(cell_of[i]), r):
            sx, sy, sz = c % m[0], (c // m[0]) % m[1], c // (m[0] * m[1])
            for s, c0 in ((sx, cx[i]), (sy, cy[i]), (sz, cz[i])):
                assert ((c0 >> 2) << 2) - 1 <= s <= ((c0 >> 2) << 2) + 4
