#!/usr/bin/env python
"""bench.py -- particle-steps/s of the DPD time step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # the CUDA path (through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...   # stock CPU LAMMPS of the reference tree

Workload (default): BASELINE configs[3] -- the rho = 4 DPD fluid of example/simple/sp.run in a 200^3 box (32,000,000
particles), spatially decomposed over the N GPUs ("scaling": "strong"; N = 1 holds the whole box: 20.5 GB of table).
`--case C` selects the round-1 weak-scaling mode instead (one C^3 brick per GPU; `--case 64` = configs[1],
`--case 48 --precision dp` = configs[2]).  A "step" is one DPD time step: velocity-Verlet halves + halo refresh + pair
force; every 5th step also wrap + migration + 2-level reorder + ghost rebuild + neighbor-list build.
`--workload polymer_channel` times BASELINE configs[4] instead (amphiphilic bead-spring chains in a driven channel, cubic --box,
default 64).  Besides the contract's keys the line carries `e2e_lmp` (N = 1: sp.run case 64 through the LAMMPS binary
lmp_meso_b200, LAMMPS' own loop time), `phases` (event-timed phases of the library), `parity_check` (small-box comparison with
the oracle on the same processor grid, run before anything is timed) and `gpu_launches` (the library's own launch count).
Prints ONE JSON line on stdout.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BOX = 200                      # box edge of the default (north_star) workload
CASE = 64                      # box edge of the cpu_baseline sample and of `--case` runs
RHO = 4


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled through NVML every ms; only samples that fall inside the timed region
    (mark_begin .. mark_end) are reported.  Falls back to one nvidia-smi query if NVML cannot be loaded."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index, period=0.001):
        super().__init__(daemon=True)
        self.index, self.period, self.stop_flag, self.rows = index, period, False, []
        self.t0 = self.t1 = None
        self.h = self.nv = None
        self.sm_max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else index
            self.h, self.nv = pynvml.nvmlDeviceGetHandleByIndex(phys), pynvml
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def run(self):
        if self.h is None:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                self.rows.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                  int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))))
            except Exception:
                try:
                    self.rows.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)), 0))
                except Exception:
                    break
            time.sleep(self.period)

    def mark_begin(self): self.t0 = time.perf_counter()
    def mark_end(self): self.t1 = time.perf_counter()

    def summary(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1.0)
        if self.h is None:
            return self._smi_once()
        rows = [r for r in self.rows if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or r[0])] or self.rows[-3:]
        sm = [r[1] for r in rows]
        bits = 0
        for r in rows:
            bits |= r[2]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max,
                "reasons": [n for n, b in self.REASONS if bits & b], "samples": len(sm), "source": "nvml, %g ms period, timed region only" % (1e3 * self.period)}

    def _smi_once(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=20).stdout.strip().split("\n")[0]
            r = [t.strip() for t in out.split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(r[0]), "sm_max_mhz": float(r[1]), "reasons": [n for i, n in enumerate(names) if r[2 + i].lower().startswith("active")],
                    "samples": 1, "source": "nvidia-smi after the timed region (NVML unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "unavailable"}


def procgrid_for(n):
    """LAMMPS ProcMap: the factorisation with the smallest brick surface, first found wins (cubic box)."""
    best, bs = (1, 1, n), None
    for px in range(1, n + 1):
        if n % px:
            continue
        for py in range(1, n // px + 1):
            if (n // px) % py:
                continue
            pz = n // px // py
            s = 1.0 / (px * py) + 1.0 / (px * pz) + 1.0 / (py * pz)
            if bs is None or s < bs - 1e-12:
                best, bs = (px, py, pz), s
    return best


def workload_of(args, world):
    """(box dims, processor grid, brick dims, scaling, description).  Default: the 200^3 box split over the ranks;
    --case C: one C^3 brick per rank."""
    grid = procgrid_for(world)
    if args.case is not None:
        L = args.case
        dims = tuple(g * L for g in grid)
        return dims, grid, (L, L, L), "weak", "example/simple %s.run case=%d per GPU" % (args.precision, L)
    B = args.box
    if any(B % g for g in grid):
        raise SystemExit("--box %d is not divisible by the processor grid %s" % (B, grid))
    return (B, B, B), grid, tuple(B // g for g in grid), "strong", "DPD fluid %d^3 box (example/simple %s.run deck) decomposed over %d GPU(s)" % (B, args.precision, world)


# ---------------------------------------------------------------------------------------- reference arm
STOCK_DECK = ("dimension 3\nunits lj\natom_style atomic\ncommunicate single vel yes\nneighbor 0.3 bin\n"
              "neigh_modify delay 0 every 5 check no\nread_data c.data\npair_style dpd 1.0 1.0 419084618\n"
              "pair_coeff 1 1 15 4.5 1.0\nvelocity all create 1.0 788662042 loop all\nfix 3 all nve\n"
              "thermo_style custom step temp press cpu spcpu\nthermo 100\ntimestep 0.005\nrun %d\nrun %d\n")


def host_copies(L):
    """How many serial LAMMPS processes the CPU legs run side by side: the image has no MPI, so the closest stand-in for
    `mpirun -np P` on the box's own cores is P independent copies, each simulating one brick of the P-way decomposition
    of the L^3 box as its own periodic system (same particles per core, no halo traffic: an upper bound of the MPI run)."""
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    p = 1
    while p * 2 <= min(cores, 64) and all(L % g == 0 and L // g >= 4 for g in procgrid_for(p * 2)):
        p *= 2
    return p


def _write_brick(args):
    from meso_b200 import workload
    path, brick, seed = args
    workload.write_data(path, workload.dpd_fluid(brick if len(set(brick)) > 1 else brick[0], seed=seed), brick)
    return path


def stock_lammps_rate(L, warm, steps):
    """particle-steps/s of stock pair_style dpd + fix nve on the L^3 box spread over host_copies(L) serial processes
    (returns rate, copies, description) or None when oracle/_ref/lmp_serial did not travel"""
    from meso_b200 import workload
    lmp = os.path.join(ROOT, "oracle", "_ref", "lmp_serial")
    if not os.path.exists(lmp):
        return None
    p = host_copies(L)
    grid = procgrid_for(p)
    brick = tuple(L // g for g in grid)
    with tempfile.TemporaryDirectory() as d:
        procs, jobs = [], []
        for c in range(p):
            dc = os.path.join(d, "c%d" % c)
            os.mkdir(dc)
            jobs.append((os.path.join(dc, "c.data"), brick, workload.DEFAULT_SEED + c))
            open(os.path.join(dc, "in.ref"), "w").write(STOCK_DECK % (warm, steps))
        if p > 1 and brick[0] * brick[1] * brick[2] * RHO > 50000:
            import multiprocessing as mp                    # the data files are written by the same cores that run them afterwards
            with mp.get_context("fork").Pool(p) as pool:
                pool.map(_write_brick, jobs)
        else:
            for j in jobs:
                _write_brick(j)
        for c in range(p):
            dc = os.path.join(d, "c%d" % c)
            procs.append(subprocess.Popen([lmp, "-meso", "off", "-in", "in.ref", "-log", "none"], cwd=dc, stdout=subprocess.PIPE,
                                          stderr=subprocess.STDOUT, text=True))
        loops = []
        for pr in procs:
            out = pr.communicate()[0]
            t = [float(s.split()[3]) for s in out.split("\n") if s.startswith("Loop time of")]
            if pr.returncode != 0 or len(t) < 2:
                return ("stock LAMMPS run failed: " + out[-200:].replace("\n", " "), 0, "")
            loops.append(t[-1])
    n = RHO * L ** 3
    what = ("stock LAMMPS 30Sep2013 pair_style dpd + fix nve (oracle/_ref/lmp_serial): %d serial processes side by side, each one %s brick "
            "of the %d^3 box as its own periodic system (no MPI in the image: stand-in for mpirun -np %d without halo traffic), "
            "slowest process counts" % (p, "x".join(str(b) for b in brick), L, p))
    return (n * steps / max(loops), p, what)


def reference_arm(args):
    """Stock LAMMPS pair_style dpd + fix nve on the box's host cores, on this arm's workload:
    the whole 200^3 box (or the --case brick set) split over the host cores, a bounded number of time steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from meso_b200 import workload
    world = max(1, args.gpus)
    dims, grid, brick, scaling, wl = workload_of(args, world)
    if len(set(dims)) != 1:
        L = round((dims[0] * dims[1] * dims[2]) ** (1.0 / 3.0))           # weak mode on a non-cubic grid: the cube of equal volume
        while L ** 3 < dims[0] * dims[1] * dims[2]:
            L += 1
    else:
        L = dims[0]
    n = RHO * L ** 3
    sample_steps = max(1, min(args.steps, args.ref_steps))
    warm = max(1, min(args.warmup, 3))
    r = stock_lammps_rate(L, warm, sample_steps)
    if r is not None and r[1] == 0:
        print(json.dumps({"impl": "reference", "unavailable": r[0]}))
        return
    if r is not None:
        value, cores, what = r
        kind = "reference"
    else:
        import oracle
        Ls = min(L, 32)                                      # the scalar port: a bounded sub-box of the same fluid
        n = RHO * Ls ** 3
        w = oracle.World((0, 0, 0), (Ls, Ls, Ls))
        w.set_atoms(workload.dpd_fluid(Ls), workload.maxwell_velocities(n))
        w.setup()
        w.run(warm)
        t0 = time.perf_counter()
        w.run(sample_steps)
        value = n * sample_steps / (time.perf_counter() - t0)
        kind, cores, what = "port", 1, "oracle/meso_oracle.c (scalar C restatement of the MESO algorithm) on a %d^3 sub-box, 1 core" % Ls
    line = {"impl": "reference", "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * RHO * L ** 3 / value, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %d particles, rho=4, rc=1, skin 0.3, rebuild every 5 steps, dt 0.005" % (wl, RHO * L ** 3),
                       "what": what},
            "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": kind,
                             "sample": "%d time steps after %d warm-up steps (of --steps %d); %s" % (sample_steps, warm, args.steps, what)},
            "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_baseline_sample(L):
    """bounded CPU sample for the cpu_baseline key: stock LAMMPS if it travelled, else the oracle port.  Always the
    case-64 box (particle-steps/s is size-normalised; the 200^3 box is what --impl reference times)."""
    from meso_b200 import workload
    n = RHO * L ** 3
    steps = 10
    r = stock_lammps_rate(L, 2, steps)
    if r is not None and r[1] > 0:
        return {"value": r[0], "unit": "particle-steps/s", "cores": r[1], "kind": "reference",
                "sample": "%d steps of the case=%d box (1,048,576 particles; after 2 warm-up steps); %s" % (steps, L, r[2])}
    import oracle
    Ls = min(L, 32)
    n = RHO * Ls ** 3
    w = oracle.World((0, 0, 0), (Ls, Ls, Ls))
    w.set_atoms(workload.dpd_fluid(Ls), workload.maxwell_velocities(n))
    w.setup()
    steps = 5
    t0 = time.perf_counter()
    w.run(steps)
    t = time.perf_counter() - t0
    return {"value": n * steps / t, "unit": "particle-steps/s", "cores": 1, "kind": "port",
            "sample": "oracle C port, %d steps of a %d^3 box of the same fluid" % (steps, Ls)}


# ---------------------------------------------------------------------------------------- the real plugin path
LMP_DECK = ("dimension 3\nunits lj\natom_style dpd/atomic/meso\nneighbor 0.3 bin\nneigh_modify delay 0 every 5 check no\n"
            "read_data ${case}.data\nrun_style mvv/meso\npair_style %s 1.0 419084618\npair_coeff 1 1 15 4.5 3.0 1.0 1.0\n"
            "compute mythermo all temp/meso\nvelocity all create 1.0 788662042 loop all\nfix 3 all nve/meso\n"
            "thermo_style custom step temp cpu spcpu\nthermo 100\nthermo_modify temp mythermo\ntimestep 0.005\nrun %d\n")


def lammps_deck_rate(L, precision, steps):
    """example/simple/{sp,dp}.run (case L) through lammps/_build/lmp_meso_b200 -- the reference's own LAMMPS core with this
    repository's package -- on GPU 0: particle-steps/s from LAMMPS' own `Loop time` line (host arrays in, thermo every 100 steps).
    None when the binary did not travel."""
    from meso_b200 import workload
    lmp = os.path.join(ROOT, "lammps", "_build", "lmp_meso_b200")
    if not os.path.exists(lmp):
        return None
    with tempfile.TemporaryDirectory() as d:
        workload.write_data(os.path.join(d, "%d.data" % L), workload.dpd_fluid(L), L)
        open(os.path.join(d, "in.run"), "w").write(LMP_DECK % ("dpd/fast/meso" if precision == "sp" else "dpd/meso", steps))
        out = subprocess.run([lmp, "-in", "in.run", "-var", "case", str(L), "-log", "none"], cwd=d, capture_output=True, text=True, timeout=900)
    t = [float(l.split()[3]) for l in out.stdout.split("\n") if l.startswith("Loop time of")]
    if out.returncode != 0 or not t:
        return {"error": (out.stdout[-200:] + out.stderr[-200:]).replace("\n", " ")}
    n = RHO * L ** 3
    return {"value": n * steps / t[-1], "unit": "particle-steps/s", "loop_time_s": t[-1], "steps": steps,
            "what": "lmp_meso_b200 -in %s.run -var case %d (%d particles): LAMMPS' own Loop time for `run %d` with thermo 100 "
                    "(temp/meso reduces on the device: no per-atom download at thermo steps)" % (precision, L, n, steps)}


# ---------------------------------------------------------------------------------------- parity preflight
def parity_preflight(rank, world, grid, local_rank, dist):
    """Small-box check of the path that is about to be timed against the CPU oracle (the checker, not the product):
    one rank = __graft_entry__.smoke()'s comparison; N ranks = tests/mgpu_check.py at L = 12 on the SAME processor grid
    (fp64 lockstep over 12 steps incl. migration, fp32 forces, ordered ghosts + neighbor lists + TEA signatures)."""
    t0 = time.perf_counter()
    out = {"ok": False, "ranks": world, "grid": list(grid)}
    try:
        if world == 1:
            import oracle
            from meso_b200 import workload
            from meso_b200.engine import dpd_fluid_deck
            L = 8
            x = workload.dpd_fluid(L)
            v = workload.maxwell_velocities(len(x))
            errs = {}
            for precision, tol in (("sp", 1e-5), ("dp", 1e-12)):
                m = dpd_fluid_deck(L, precision, device=local_rank, x=x, v=v)
                m.setup()
                w = oracle.World((0, 0, 0), (L, L, L), precision=0 if precision == "sp" else 1)
                w.set_atoms(x, v)
                w.setup()
                cg, rg = m.neighbors()
                co, ro = w.neighbors()
                mask = np.arange(ro.shape[1])[None, :] < co[:, None]
                assert np.array_equal(cg, co) and np.array_equal(rg[mask], ro[mask]), "neighbor lists differ from the oracle"
                assert np.array_equal(m.packed()[1].view(np.uint32), w.packed()[1].view(np.uint32)), "TEA signatures differ"
                fg, fo = m.download(("f",))["f"], w.atoms()["f"]
                mag = np.linalg.norm(fo, axis=1)
                errs[precision] = float((np.linalg.norm(fg - fo, axis=1) / np.maximum(mag, mag.mean())).max())
                assert errs[precision] <= tol, (precision, errs[precision])
                m.close()
            out.update(ok=True, box="8^3", force_err_sp=errs["sp"], force_err_dp=errs["dp"], lists="bit-exact", signatures="bit-exact")
        else:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import mgpu_check
            from meso_b200.engine import Meso
            dims = (12, 12, 12)
            inp = mgpu_check.make_inputs(dims)
            errs = {}
            for precision in ("dp", "sp"):
                ids = [Meso.unique_id() if rank == 0 else None]
                dist.broadcast_object_list(ids, src=0)
                errs[precision] = mgpu_check.check(precision, rank, grid, local_rank, dims, inp, ids[0], steps=12)
            flags = [None] * world
            dist.all_gather_object(flags, errs)
            out.update(ok=True, box="12^3", force_err_sp=max(f["sp"] for f in flags), force_err_dp=max(f["dp"] for f in flags),
                       lists="bit-exact (ordered, every rank)", ghosts="bit-exact order and values", dp_lockstep_steps=12)
    except Exception as e:                                   # a failed preflight is reported, and the bench line says so
        out["error"] = "%s: %s" % (type(e).__name__, str(e)[:300])
        if world > 1:
            raise
    out["seconds"] = round(time.perf_counter() - t0, 2)
    return out


# ---------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="meso_b200", choices=["meso_b200", "reference"])
    ap.add_argument("--box", type=int, default=BOX, help="edge of the periodic box that is decomposed over the GPUs (200 = BASELINE configs[3])")
    ap.add_argument("--case", type=int, default=None, help="weak-scaling mode: box edge per GPU brick (64 = BASELINE configs[1])")
    ap.add_argument("--workload", default="fluid", choices=["fluid", "polymer_channel"],
                    help="polymer_channel = BASELINE configs[4]: amphiphilic bead-spring chains (bond_style harmonic, 1-2 exclusions, three atom "
                         "types) in solvent between walls across z, driven by pois/meso; a cubic --box (default 64) decomposed over the GPUs")
    ap.add_argument("--precision", default="sp", choices=["sp", "dp"])
    ap.add_argument("--ref-steps", type=int, default=20, help="time steps the reference arm actually runs (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity preflight")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--thermo", type=int, default=100, help="e2e leg: thermo/output interval (deck: 100)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    # stdout carries exactly one JSON line: keep a private handle on it and point fd 1 at stderr, so that banners printed by
    # native libraries (NCCL prints its version on this pool's boxes) cannot end up in front of the JSON
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from meso_b200 import workload
    from meso_b200.engine import Meso

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    channel = args.workload == "polymer_channel"
    if channel:
        if args.case is not None:
            raise SystemExit("--workload polymer_channel takes --box (a cubic box decomposed over the GPUs), not --case")
        if args.box == BOX:
            args.box = 64
    dims, grid, brick, scaling, wl = workload_of(args, world)
    if channel:
        wl = "amphiphilic bead-spring chains in a driven channel, %d^3 box (walls across z) decomposed over %d GPU(s)" % (args.box, world)
    parity = None if args.no_parity else parity_preflight(rank, world, grid, local_rank, dist)

    # every rank generates its own brick of the box (4 uniformly random points per unit cell, cells x-fastest: the layout of
    # example/simple/25.data); tags are consecutive per brick
    nloc = RHO * brick[0] * brick[1] * brick[2]
    nglob = nloc * world
    loc = (rank // (grid[1] * grid[2]), (rank // grid[2]) % grid[1], rank % grid[2])
    typ = bonds = None
    if channel:
        # every rank builds the whole channel from the same seed and keeps the beads of its brick (chains cross brick faces)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from mgpu_check import AMPHI_COEFF
        xa, ta, ga, nb, bt, ba = workload.amphiphilic_channel(args.box)
        va = workload.maxwell_velocities(len(xa), seed=99)
        lo = np.array([loc[d] * brick[d] for d in range(3)], dtype=np.float64)
        mine = np.all((xa >= lo) & (xa < lo + np.array(brick, dtype=np.float64)), axis=1)
        x, v, tag, typ = np.ascontiguousarray(xa[mine]), np.ascontiguousarray(va[mine]), ga[mine].astype(np.int32), ta[mine].astype(np.int32)
        bonds = (np.ascontiguousarray(nb[mine]), np.ascontiguousarray(bt[mine]), np.ascontiguousarray(ba[mine]), len(xa))
        nglob, nloc = len(xa), int(mine.sum())
        del xa, va, ta, ga, nb, bt, ba
    else:
        x = workload.dpd_fluid(brick if len(set(brick)) > 1 else brick[0], seed=workload.DEFAULT_SEED + rank)
        x += np.array([loc[d] * brick[d] for d in range(3)], dtype=np.float64)
        v = workload.maxwell_velocities(nloc, seed=788662042 + rank)
        tag = (np.arange(nloc, dtype=np.int64) + 1 + rank * nloc).astype(np.int32)

    def deck():
        m = Meso(local_rank)
        m.box((0.0, 0.0, 0.0), dims, (1, 1, 0) if channel else (1, 1, 1))
        if world > 1:
            ids = [Meso.unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            m.decomposition(rank, grid, ids[0])
        m.masses([0.0, 1.0, 1.0, 1.0] if channel else [0.0, 1.0])
        m.neighbor(0.3, "bin")
        m.neigh_modify(delay=0, every=5, check=False)
        m.pair_style("dpd/fast/meso" if args.precision == "sp" else "dpd/meso", 1.0, 419084618)
        if channel:
            for (a, b), a0 in AMPHI_COEFF.items():
                m.pair_coeff(a, b, a0, 4.5, 3.0, 1.0, 1.0)
        else:
            m.pair_coeff(1, 1, 15, 4.5, 3.0, 1.0, 1.0)
        m.timestep(0.005)
        return m

    def load(m):
        """the deck's read_data: atoms (pinned host buffers -> device), and for the channel the bond table and the fixes"""
        m.upload(xp.numpy(), vp.numpy(), tag=tp.numpy(), type=typ)
        if channel:
            m.bond_style("harmonic/meso", 1)
            m.bond_coeff(1, 50.0, 0.5)
            m.special_bonds(0.0)
            m.bonds(bonds[0], bonds[1], bonds[2], tag_max=bonds[3])
            m.unfix_all()
            m.fix("solid_bound/meso", "z", "rho5rc1s1"); m.fix("pois/meso", "z", "x", 0.2)

    xp = torch.from_numpy(x).pin_memory()
    vp = torch.from_numpy(v).pin_memory()
    tp = torch.from_numpy(tag).pin_memory()
    del x, v, tag

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg: `value`
    warmup = max(args.warmup, 3)
    m = deck()
    load(m)
    m.setup()
    stream = torch.cuda.ExternalStream(m.stream())
    m.run(warmup)
    m.sync()
    n_bar = float(m.pair_count().mean())
    m.timers(enable=True, reset=True)
    m.launch_count(reset=True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    e0.record(stream)
    m.run(args.steps)
    e1.record(stream)
    m.sync()
    sampler.mark_end()
    barrier()
    launches = m.launch_count()
    ms = e0.elapsed_time(e1)
    clocks = sampler.summary()
    tm = m.timers(enable=False)
    T_end = m.temperature()
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = nglob * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (the pair-force kernel)
    peak, peak_src = peaks()
    once = os.environ.get("MESO_PAIR_ONCE", "1") != "0"
    pair_ms, pair_calls = tm["pair"]
    b_force = 4.0 * n_bar + 32.0 + 24.0                        # algorithmic bytes per particle per force evaluation
    # per rank: every step evaluates the force on all nloc particles (one launch on 1 GPU; bulk + border launches on N > 1)
    achieved = b_force * nloc * args.steps / (pair_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tp_path = os.path.join(ROOT, "profiles", "force_kernel_traffic.json")
    if os.path.exists(tp_path):
        tj = json.load(open(tp_path))
        key = "%s%s" % (args.precision, "_once" if once else "")
        if "dram_bytes_per_particle_" + key in tj:
            # ncu --set full capture of this kernel (profiles/): bytes per particle of that capture x particles per launch here
            traffic = tj["dram_bytes_per_particle_" + key] * nloc * args.steps / max(pair_calls, 1)
            traffic_src = tj.get("source_" + key)
    kname = ("k_dpd_once<%s> (each local pair once, REDG scatter)" if once else "k_dpd<%s,0> (two-sided, +fused final_integrate)") % \
        ("float" if args.precision == "sp" else "double")
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_particle": b_force, "particles_per_launch": nloc * args.steps / max(pair_calls, 1),
                "avg_launch_ms": pair_ms / max(pair_calls, 1), "share_of_step": pair_ms / ms,
                "note": "not HBM-bound: a list-based gather kernel that waits for gather latency (ncu, profiles/r02_s3_force_*: issue slots "
                        "52 %, L1 data pipe 55 %, DRAM 16 %, 0.8x the algorithmic bytes moved: the owned prefix is half a row); see DESIGN.md s3.1"}
    phases = {k: {"ms_total": round(v[0], 3), "calls": int(v[1])} for k, v in tm.items()}
    m.close()

    # ---- end-to-end leg through the C ABI with HOST buffers: the deck's `run K` with `thermo 100`
    # upload of all atoms (H2D from pinned memory), setup, K steps, temp/meso every `thermo` steps, and the positions,
    # velocities, forces and tags of every atom read back at the end of the run (transfer_pre_output)
    e2e = None
    if not args.no_e2e:
        m = deck()
        ncap = nloc if world == 1 else nloc + nloc // 8 + 1024      # atoms migrate between bricks: a rank's count drifts around nloc
        out = {k: torch.empty((ncap, 3), dtype=torch.float64).pin_memory() for k in ("x", "v", "f")}
        otag = torch.empty(ncap, dtype=torch.int32).pin_memory()
        import ctypes as C
        vp_ = lambda t_: C.c_void_p(t_.data_ptr())

        def run_deck(nsteps):
            t0 = time.perf_counter()
            load(m)
            m.ntimestep = 0
            m.setup()
            t_setup = time.perf_counter() - t0                   # meso_setup returns after the device finished (it reads the counts back)
            done, temps = 0, []
            while done < nsteps:
                chunk = min(args.thermo, nsteps - done)
                m.run(chunk)
                done += chunk
                temps.append(m.temperature())
            t1 = time.perf_counter()
            m._chk(m.L.meso_atoms_download(m.h, ncap, vp_(out["x"]), vp_(out["v"]), vp_(out["f"]), vp_(otag), None, None, None))
            t_down = time.perf_counter() - t1
            d2h = m.counts()["nlocal"] * (72 + 4) + 16 * len(temps)
            return d2h, temps, t_setup, t_down

        run_deck(min(args.steps, 10))                        # warm-up (allocations, first-touch)
        barrier()
        t0 = time.perf_counter()
        d2h, temps, t_setup, t_down = run_deck(args.steps)
        barrier()
        wall = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([wall, t_setup, t_down], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall, t_setup, t_down = (float(a) for a in t.tolist())
        h2d = nloc * (48 + 4) + (nloc * (4 + 4 + 4 * int(bonds[1].size // max(len(bonds[0]), 1)) * 2) if channel else 0)
        e2e = {"value": nglob * args.steps / wall, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d / args.steps,
               "d2h_bytes_per_step": d2h / args.steps, "setup_ms": 1e3 * t_setup, "download_ms": 1e3 * t_down,
               "stepping_ms": 1e3 * (wall - t_setup - t_down), "wall_ms": 1e3 * wall,
               "what": "C-ABI deck run with pinned HOST buffers: meso_atoms_upload (H2D) + meso_setup (first reorder, ghosts, neighbor table, "
                       "setup force; = setup_ms) + %d step(s) in chunks of %d each followed by temp/meso (D2H scalar) + one meso_atoms_download "
                       "of x,v,f,tag (D2H; = download_ms); wall clock over everything, max over ranks" % (args.steps, args.thermo)}
        m.close()

    line = {"metric": "particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f32" if args.precision == "sp" else "f64", "data": "synthetic",
            "config": {"workload": "%s: rho=4, %d particles (%d per GPU), box %dx%dx%d, rc=1, skin 0.3, rebuild every 5 steps, dt 0.005"
                                   % (wl, nglob, nloc, *dims),
                       "procgrid": list(grid), "brick": list(brick), "pair_style": "dpd/fast/meso" if args.precision == "sp" else "dpd/meso",
                       **({"bond_style": "harmonic/meso", "special_bonds": "lj 0 1 1", "atom_types": 3,
                           "fixes": ["solid_bound/meso z rho5rc1s1", "pois/meso z x 0.2", "nve/meso"]} if channel else {}),
                       "l2": "per-step working set (>= 0.8 GB per million particles) exceeds the 126 MB L2; no explicit flush",
                       "mean_neighbors": n_bar, "temperature_end": T_end},
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "phases": phases, "clocks": clocks, "parity_check": parity}
    if rank == 0 and world == 1 and not args.no_e2e and not channel:
        line["e2e_lmp"] = lammps_deck_rate(CASE, args.precision, 1000)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(CASE)
    if rank == 0:
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        # a rank that dies inside a collective would leave its peers waiting for the launcher's timeout: leave at once
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
