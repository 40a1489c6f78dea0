"""debug probe: gang of N bricks on GPU 0, sp run, prints the temperature trace (python tools/gang_probe.py N [steps])"""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import mgpu_check
from meso_b200.engine import Meso
n = int(sys.argv[1]); steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
prec = sys.argv[3] if len(sys.argv) > 3 else "sp"
dims = (12, 12, 12)
inp = mgpu_check.make_inputs(dims, False)
m = Meso([0] * n)
m.box((0.0, 0.0, 0.0), dims, (1, 1, 1)); m.masses([0.0, 1.0]); m.neighbor(0.3, "bin"); m.neigh_modify(delay=0, every=5, check=False)
m.pair_style("dpd/fast/meso" if prec == "sp" else "dpd/meso", 1.0, 419084618); m.pair_coeff(1, 1, 15, 4.5, 3.0, 1.0, 1.0); m.timestep(0.005)
m.upload(inp["x"], inp["v"], tag=inp["tag"], type=inp["typ"])
m.setup()
out = [m.temperature()]
for s in range(steps):
    m.run(1)
    out.append(m.temperature())
print(n, prec, os.environ.get("MESO_PAIR_ONCE"), os.environ.get("MESO_PAIR_TEX"), " ".join("%.4g" % t for t in out), flush=True)
m.close()
