"""Several bricks behind ONE handle (meso_create_gang, meso_b200/csrc/gang.cu): the single-process counterpart of one MPI rank
per GPU.  The gang deals the atoms out by brick, drives every brick on its own host thread and sums the reductions; the result
must be the oracle's simulated-rank world of the same processor grid -- brick after brick, bit for bit where the work is integer.
One brick per GPU: the cases need as many GPUs as bricks and are skipped on smaller boxes (a device MAY be listed more than once,
but kernels of one CUDA context are not time-sliced against each other the way processes are: a halo kernel that waits for its
neighbor's flag can starve the neighbor -- measured in round 2: 2-8 bricks on one B200 finish or time out from run to run.  The
multi-rank path itself is covered on a single GPU by tests/test_multi_gpu.py, one process per rank)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mgpu_check  # noqa: E402
import oracle  # noqa: E402
from meso_b200 import lib as _lib  # noqa: E402
from meso_b200.engine import Meso, MesoError  # noqa: E402

pytestmark = pytest.mark.gpu

GRIDS = {2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}


def devices_for(nbrick):
    n = _lib.load().meso_device_count()
    if n < nbrick:
        pytest.skip("needs %d GPUs (one brick per device)" % nbrick)
    return list(range(nbrick))


def deck(m, dims, ntypes, polymer, precision, channel=False):
    m.box((0.0, 0.0, 0.0), dims, (1, 1, 0) if channel else (1, 1, 1))
    m.masses([0.0] + [1.0] * ntypes)
    m.neighbor(0.3, "bin")
    m.neigh_modify(delay=0, every=5, check=False)
    m.pair_style("dpd/fast/meso" if precision == "sp" else "dpd/meso", 1.0, 419084618)
    if channel:
        for (a, b), a0 in mgpu_check.AMPHI_COEFF.items():
            m.pair_coeff(a, b, a0, 4.5, 3.0, 1.0, 1.0)
    elif polymer:
        m.pair_coeff(1, 1, 25, 4.5, 3.0, 1.0, 1.0); m.pair_coeff(1, 2, 40, 4.5, 3.0, 1.0, 1.0); m.pair_coeff(2, 2, 25, 4.5, 3.0, 1.0, 1.0)
    else:
        m.pair_coeff(1, 1, 15, 4.5, 3.0, 1.0, 1.0)
    m.timestep(0.005)


def oracle_world(dims, grid, inp, precision):
    channel = inp["channel"]
    w = oracle.World((0, 0, 0), dims, periodic=(1, 1, 0) if channel else (1, 1, 1), procgrid=grid, precision=1 if precision == "dp" else 0,
                     ntypes=inp["ntypes"], coeff=inp["coeff"], **({"mass": [0.0, 1.0, 1.0, 1.0]} if channel else {}))
    w.set_atoms(inp["x"], inp["v"], tag=inp["tag"], type=inp["typ"])
    if inp["polymer"]:
        nbond, btype, batom = inp["bonds"]
        w.set_bonds(nbond, btype, batom, tag=inp["tag"], k=[0.0, 50.0], r0=[0.0, 0.5], special_lj12=0.0)
    if channel:
        w.fix_solid_bound("z"); w.fix_pois(2, 0, 0.2)
    return w


def step_by_phases(m, polymer):
    B, O, LOC = _lib.MESO_BULK, _lib.MESO_BORDER, _lib.MESO_LOCAL
    m.ntimestep = m.ntimestep + 1
    m.initial_integrate()
    if m.neighbor_decide():
        m.rebuild()
        m.force_clear(LOC)
        m.pair_compute(B)
    else:
        m.force_clear(B)
        m.pair_compute(B)
        m.forward_comm()
        m.force_clear(O)
    m.pair_compute(O)
    if polymer:
        m.bond_compute()
    m.final_integrate()


@pytest.mark.parametrize("nbrick,L,polymer,phases", [(2, 12, False, False), (4, 12, False, False), (2, 10, True, False), (2, 12, False, True),
                                                     (4, 12, True, False), (8, 12, False, False), (4, 8, "channel", False),
                                                     (8, 8, "channel", False)])
def test_gang_equals_the_oracle_world_brick_after_brick(nbrick, L, polymer, phases):
    """polymer == "channel": BASELINE configs[4] -- amphiphilic chains between walls on a decomposed non-periodic dimension"""
    dims, grid = (L, L, L), GRIDS[nbrick]
    channel = polymer == "channel"
    polymer = bool(polymer)
    inp = mgpu_check.make_inputs(dims, polymer, channel)
    for precision in ("dp", "sp"):
        w = oracle_world(dims, grid, inp, precision)
        w.setup(eflag=1, vflag=1)
        m = Meso(devices_for(nbrick))
        assert m.L.meso_gang_size(m.h) == nbrick
        deck(m, dims, inp["ntypes"], polymer, precision, channel)
        m.upload(inp["x"], inp["v"], tag=inp["tag"], type=inp["typ"])
        if polymer:
            nbond, btype, batom = inp["bonds"]
            m.bond_style("harmonic/meso", 1)
            m.bond_coeff(1, 50.0, 0.5)
            m.special_bonds(0.0)
            m.bonds(nbond, btype, batom, tag_max=len(inp["x"]))
        if channel:
            m.fix("solid_bound/meso", "z", "rho5rc1s1"); m.fix("pois/meso", "z", "x", 0.2)
        m.setup(eflag=1, vflag=1)
        parts = [w.atoms(r) for r in range(nbrick)]
        cat = lambda key: np.concatenate([a[key][:a["nlocal"]] for a in parts])
        cg = m.counts()
        assert cg["nlocal"] == len(inp["x"]) and cg["nghost"] == sum(w.counts(r)["nghost"] for r in range(nbrick))
        ag = m.download()
        assert np.array_equal(ag["tag"], cat("tag")), "order of the local atoms differs (brick after brick, sorted inside)"
        assert np.array_equal(ag["x"], cat("x"))
        fo = cat("f")
        mag = np.linalg.norm(fo, axis=1)
        err = (np.linalg.norm(ag["f"] - fo, axis=1) / np.maximum(mag, mag.mean())).max()
        assert err <= ((2e-5 if polymer else 1e-5) if precision == "sp" else (1e-11 if polymer else 1e-12)), err
        assert abs(m.temperature() - w.temperature()) < 1e-12
        if polymer:
            assert abs(m.bond_energy() - w.bond_energy()) < 1e-10 * abs(w.bond_energy())
        if precision == "dp":
            for _ in range(12):
                if phases:
                    step_by_phases(m, polymer)
            if not phases:
                m.run(12)
            w.run(12)
            ag = m.download()
            parts = [w.atoms(r) for r in range(nbrick)]
            n_per = [a["nlocal"] for a in parts]
            assert m.counts()["nlocal"] == sum(n_per)
            og, oo = np.argsort(ag["tag"]), np.argsort(cat("tag"))
            assert np.abs(ag["x"][og] - cat("x")[oo]).max() < 1e-10 and np.abs(ag["v"][og] - cat("v")[oo]).max() < 1e-10
            # ownership after the migrations of steps 5 and 10: the download is brick after brick
            start = 0
            for a in parts:
                assert set(ag["tag"][start:start + a["nlocal"]].tolist()) == set(a["tag"][:a["nlocal"]].tolist())
                start += a["nlocal"]
            assert abs(m.temperature() - w.temperature()) < 1e-10
        else:
            m.run(12)
            assert 0.5 < m.temperature() < (6.0 if polymer else 3.0)
        m.close()


def test_a_device_listed_twice_is_refused_up_front():
    """one brick per device: bricks of one CUDA context can starve each other (see the module docstring); the library says so at
    creation instead of timing out in some later step"""
    with pytest.raises(MesoError, match="listed more than once"):
        Meso([0, 0])


def test_gang_of_one_is_a_plain_context():
    m = Meso([0])
    assert m.L.meso_gang_size(m.h) == 1
    m.close()


def test_gang_refuses_per_brick_exports_and_explicit_decomposition(monkeypatch):
    monkeypatch.setenv("MESO_GANG_SHARE_DEVICE", "1")       # creation and settings only: nothing runs on the shared device
    m = Meso([0, 0])
    m.box((0.0, 0.0, 0.0), (8.0, 8.0, 8.0), (1, 1, 1))
    with pytest.raises(MesoError):
        m.decomposition(0, (1, 1, 2))
    with pytest.raises(MesoError):
        m.pair_count()
    m.close()
