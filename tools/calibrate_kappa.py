"""Developer tool (GPU): measure H / DH error of the tcgen05 engine vs the fp64 oracle for the
current FPS_TC_SEG / FPS_TC_KAPPA / FPS_TC_PAIR environment.  Prints one line."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_poisson_solver_b200 import Solver
from oracle import fps_oracle as o
os.makedirs("/tmp/fps_cal", exist_ok=True); os.chdir("/tmp/fps_cal")
n = int(os.environ.get("CAL_N", "3000"))
ref = "/tmp/fps_cal/ref_%d.npz" % n
rng = np.random.default_rng(11)
x, y = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
if os.path.exists(ref):
    z = np.load(ref); Hr, DHr = z["H"], z["DH"]
else:
    Hr, DHr = o.features_forward(x, y, o.init_weights(0), np.float64)
    np.savez(ref, H=Hr, DH=DHr)
s = Solver(use_weights=False, engine=os.environ.get("CAL_ENGINE", "tcgen05"))
H, DH = s._features(torch.tensor(x, device="cuda"), torch.tensor(y, device="cuda"), True)
H, DH = H[:, :700].double().cpu().numpy(), DH[:, :700].double().cpu().numpy()
def stats(a, b):
    rel = o.rel_l2(a, b)
    m = np.abs(b) > 0.1 * np.abs(b).mean()
    bias = np.mean((np.abs(a[m]) - np.abs(b[m])) / np.abs(b[m]))   # signed magnitude bias
    return rel, bias
print("seg=%s kappa=%s pair=%s | H rel %.3e bias %+.3e | DH rel %.3e bias %+.3e" % (
    os.environ.get("FPS_TC_SEG"), os.environ.get("FPS_TC_KAPPA"), os.environ.get("FPS_TC_PAIR"), *stats(H, Hr), *stats(DH, DHr)))
