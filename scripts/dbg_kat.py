"""Debug: repeated device-resident solves with / without a concurrent torch kernel on another stream."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fdm_b200
from fdm_b200 import selfcheck as G
mode = sys.argv[1]
nz, ny, nx = (int(v) for v in sys.argv[2:5])
dev = torch.device("cuda", 0)
L = fdm_b200.lib()
d = 1.0 / 1023
S = fdm_b200.LaplCube(d, d, d, d * (nx + 1), d * (ny + 1), d * (nz + 1), nx, ny, nz)
rhs = torch.rand(nz * ny * nx, dtype=torch.float64, device=dev) - 0.5
ans = torch.full_like(rhs, float("nan"))
torch.cuda.synchronize()
S.solve_device(ans.data_ptr(), rhs.data_ptr())
fdm_b200.capi.check(L.fdmb_device_synchronize(), "sync1")
print("solve 1 ok", float(ans.abs().max()))
for rep in range(3):
    ans2 = torch.full_like(rhs, float("nan"))
    if mode == "sync":
        torch.cuda.synchronize()
    S.solve_device(ans2.data_ptr(), rhs.data_ptr())
    fdm_b200.capi.check(L.fdmb_device_synchronize(), "sync2")
    print("solve", rep + 2, "ok equal:", bool(torch.equal(ans, ans2)))
