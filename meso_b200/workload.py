"""Synthetic DPD-fluid inputs shaped like the reference's example/simple/*.data.

Structure verified on the reference's 25.data: rho uniformly random
points per unit cell, cells enumerated x-fastest then y then z, ids 1..N in that
order, one atom type of mass 1.  48.data/64.data are absent from the reference
tree, so every size other than 25 comes from here (PCG64, seed recorded below).
"""
import numpy as np

DEFAULT_SEED = 20140901


def dpd_fluid(L, rho=4, seed=DEFAULT_SEED, dtype=np.float64):
    """Positions (N,3) of a rho-per-unit-cell random fluid in [0,L)^3; L may be a 3-tuple."""
    Lx, Ly, Lz = (L, L, L) if np.isscalar(L) else L
    rng = np.random.Generator(np.random.PCG64(seed))
    ncell = Lx * Ly * Lz
    frac = rng.random((ncell, rho, 3))
    c = np.arange(ncell)
    origin = np.stack([c % Lx, (c // Lx) % Ly, c // (Lx * Ly)], axis=1).astype(np.float64)
    x = (origin[:, None, :] + frac).reshape(-1, 3)
    # same text round-trip as a %.9e data file, so file-based and in-memory runs agree
    return np.ascontiguousarray(np.round(x, 9), dtype=dtype)


def maxwell_velocities(n, temperature=1.0, seed=788662042, mass=1.0):
    """Gaussian velocities with zero net momentum, scaled to exactly `temperature`
    under LAMMPS' dof = 3N-3 (what `velocity all create T seed` guarantees)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    v = rng.standard_normal((n, 3))
    v -= v.mean(axis=0)
    t = mass * (v * v).sum() / (3.0 * n - 3.0)
    v *= np.sqrt(temperature / t)
    return np.ascontiguousarray(v)


def write_data(path, x, L, types=None, ntypes=1, masses=None):
    """LAMMPS data file, atom_style atomic, identical layout to example/simple/25.data."""
    Lx, Ly, Lz = (L, L, L) if np.isscalar(L) else L
    n = len(x)
    types = np.ones(n, dtype=int) if types is None else types
    masses = [1.0] * ntypes if masses is None else masses
    with open(path, "w") as f:
        f.write("LAMMPS\n\n%d atoms\n\n%d atom types\n\n" % (n, ntypes))
        f.write("0 %g xlo xhi\n0 %g ylo yhi\n0 %g zlo zhi\n\nMasses\n\n" % (Lx, Ly, Lz))
        for t, m in enumerate(masses):
            f.write("%d %f\n" % (t + 1, m))
        f.write("\nAtoms\n\n")
        for i in range(n):
            f.write("%d %d %.9e %.9e %.9e\n" % (i + 1, types[i], x[i, 0], x[i, 1], x[i, 2]))


def write_data_bond(path, x, L, types, ntypes, num_bond, bond_type, bond_atom, masses=None, nbondtypes=1):
    """LAMMPS data file for atom_style bond / dpd/bond/meso (Atoms lines `id mol type x y z`, Bonds section).  The per-atom
    table is in the newton_bond-off layout (both partners hold each bond); the file lists every bond once (a < b)."""
    Lx, Ly, Lz = (L, L, L) if np.isscalar(L) else L
    n = len(x)
    masses = [1.0] * ntypes if masses is None else masses
    bonds = []
    for i in range(n):
        for p in range(int(num_bond[i])):
            j = int(bond_atom[i][p])
            if i + 1 < j:
                bonds.append((int(bond_type[i][p]), i + 1, j))
    # molecule id: chain number for bonded beads (consecutive tags), 0 for solvent
    mol = np.zeros(n, dtype=int)
    cur = 0
    for i in range(n):
        if num_bond[i] == 0:
            continue
        if i == 0 or num_bond[i - 1] == 0 or (i + 1) not in [int(t) for t in bond_atom[i - 1][:num_bond[i - 1]]]:
            cur += 1
        mol[i] = cur
    with open(path, "w") as f:
        f.write("LAMMPS\n\n%d atoms\n%d bonds\n\n%d atom types\n%d bond types\n\n" % (n, len(bonds), ntypes, nbondtypes))
        f.write("0 %g xlo xhi\n0 %g ylo yhi\n0 %g zlo zhi\n\nMasses\n\n" % (Lx, Ly, Lz))
        for t, m in enumerate(masses):
            f.write("%d %f\n" % (t + 1, m))
        f.write("\nAtoms\n\n")
        for i in range(n):
            f.write("%d %d %d %.9e %.9e %.9e\n" % (i + 1, mol[i], types[i], x[i, 0], x[i, 1], x[i, 2]))
        f.write("\nBonds\n\n")
        for k, (t, a, b) in enumerate(bonds):
            f.write("%d %d %d %d\n" % (k + 1, t, a, b))
    return len(bonds)


def read_data(path):
    """Minimal reader for the files above: returns (x, tag, type, boxlo, boxhi, masses)."""
    with open(path) as f:
        lines = f.read().split("\n")
    n = ntypes = 0
    lo, hi = [0.0] * 3, [0.0] * 3
    i = 0
    masses = {}
    while i < len(lines):
        s = lines[i].strip()
        if s.endswith("atoms"):
            n = int(s.split()[0])
        elif s.endswith("atom types"):
            ntypes = int(s.split()[0])
        elif s.endswith("xlo xhi"):
            lo[0], hi[0] = map(float, s.split()[:2])
        elif s.endswith("ylo yhi"):
            lo[1], hi[1] = map(float, s.split()[:2])
        elif s.endswith("zlo zhi"):
            lo[2], hi[2] = map(float, s.split()[:2])
        elif s == "Masses":
            for k in range(ntypes):
                t, m = lines[i + 2 + k].split()[:2]
                masses[int(t)] = float(m)
            i += 1 + ntypes
        elif s == "Atoms":
            body = np.loadtxt(lines[i + 2:i + 2 + n])
            order = np.argsort(body[:, 0], kind="stable")
            body = body  # keep file order: LAMMPS stores atoms in file order
            x = np.ascontiguousarray(body[:, 2:5])
            return x, body[:, 0].astype(np.int32), body[:, 1].astype(np.int32), lo, hi, \
                [0.0] + [masses.get(t, 1.0) for t in range(1, ntypes + 1)]
        i += 1
    raise ValueError("no Atoms section in %s" % path)


def polymer_melt(L, chain_len=8, rho=4, bond_len=0.7, seed=20140902, solvent_frac=0.5):
    """Bead-spring chains in DPD solvent (BASELINE configs[4] flavour): `solvent_frac` of the rho*L^3 beads are free
    solvent (type 1), the rest form linear chains of `chain_len` beads (type 2) grown as random walks of step
    `bond_len` and wrapped into the box.  Returns x, type, tag and the per-atom bond table in LAMMPS' newton_bond-off
    layout (both partners hold each bond): num_bond[n], bond_type[n][2], bond_atom[n][2] (partner tags)."""
    rng = np.random.default_rng(seed)
    n = int(rho * L ** 3)
    nchain = int(n * (1.0 - solvent_frac)) // chain_len
    npoly = nchain * chain_len
    nsolv = n - npoly
    x = np.empty((n, 3))
    typ = np.ones(n, np.int32)
    x[:nsolv] = rng.random((nsolv, 3)) * L
    start = rng.random((nchain, 3)) * L
    steps = rng.normal(size=(nchain, chain_len - 1, 3))
    steps *= bond_len / np.linalg.norm(steps, axis=2, keepdims=True)
    walk = np.concatenate([start[:, None, :], start[:, None, :] + np.cumsum(steps, axis=1)], axis=1)
    x[nsolv:] = np.mod(walk.reshape(-1, 3), L)
    x = np.round(x, 9)                       # same text round trip as a %.9e data file
    x[x >= L] = 0.0
    typ[nsolv:] = 2
    tag = np.arange(1, n + 1, dtype=np.int32)
    num_bond = np.zeros(n, np.int32)
    bond_type = np.zeros((n, 2), np.int32)
    bond_atom = np.zeros((n, 2), np.int32)
    for c in range(nchain):
        base = nsolv + c * chain_len
        for b in range(chain_len - 1):
            i, j = base + b, base + b + 1
            for a, o in ((i, j), (j, i)):
                bond_type[a, num_bond[a]] = 1
                bond_atom[a, num_bond[a]] = tag[o]
                num_bond[a] += 1
    return x, typ, tag, num_bond, bond_type, bond_atom


def amphiphilic_channel(L, chain_len=8, rho=4, bond_len=0.7, seed=20140903, solvent_frac=0.6, margin=0.3):
    """BASELINE configs[4] flavour: amphiphilic bead-spring chains (first half of a chain hydrophilic = type 2, second half
    hydrophobic = type 3) in solvent (type 1) between two walls across z (periodic in x, y).  Chains are random walks
    folded back at z = margin and z = L - margin, so no bond crosses a wall.  Returns x, type, tag, num_bond, bond_type,
    bond_atom like polymer_melt (per-atom tables, both partners hold each bond)."""
    x, typ, tag, num_bond, bond_type, bond_atom = polymer_melt(L, chain_len, rho, bond_len, seed, solvent_frac)
    rng = np.random.default_rng(seed + 1)
    n = len(x)
    nchain = int(n * (1.0 - solvent_frac)) // chain_len
    nsolv = n - nchain * chain_len
    start = rng.random((nchain, 3)) * np.array([L, L, L - 2 * margin]) + np.array([0.0, 0.0, margin])
    steps = rng.normal(size=(nchain, chain_len - 1, 3))
    steps *= bond_len / np.linalg.norm(steps, axis=2, keepdims=True)
    walk = np.concatenate([start[:, None, :], start[:, None, :] + np.cumsum(steps, axis=1)], axis=1).reshape(-1, 3)
    z = walk[:, 2] - margin
    span = L - 2 * margin
    z = np.abs(np.mod(z + span, 2 * span) - span)          # triangle wave: fold into [0, span]
    walk[:, 2] = z + margin
    walk[:, :2] = np.mod(walk[:, :2], L)
    x[nsolv:] = walk
    x[:nsolv, 2] = margin + x[:nsolv, 2] * (span / L)     # solvent: squeeze into the channel
    x = np.round(x, 9)
    x[:, :2][x[:, :2] >= L] = 0.0
    pos = (np.arange(n - nsolv) % chain_len)
    typ[nsolv:] = np.where(pos < chain_len // 2, 2, 3)
    return np.ascontiguousarray(x), typ, tag, num_bond, bond_type, bond_atom
