import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "orion-sdr_b200", "python"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, orion_b200 as ob
from signals import noise_c64
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200003
x = noise_c64(n, seed=n)
g = ob.FirDecimator(2.4e6, 8, 100e3, 38400.0)
if len(sys.argv) > 2: g.set_option(ob.OPT_USE_TMA, int(sys.argv[2]))
y = g.run(x)
print("ok", y.size, float(np.abs(y).max()))
