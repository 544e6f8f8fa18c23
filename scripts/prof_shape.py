"""Per-kernel timing of one LaplCube solve of arbitrary shape: prof_shape.py nx ny nz [reps]"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fdm_b200
from fdm_b200 import capi
nx, ny, nz = map(int, sys.argv[1:4]); reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
L = fdm_b200.lib()
S = fdm_b200.LaplCube(0.1, 0.1, 0.1, 0.1 * (nx + 1), 0.1 * (ny + 1), 0.1 * (nz + 1), nx, ny, nz)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
rhs = torch.rand(nx * ny * nz, dtype=torch.float64, device="cuda") - 0.5
ans = torch.empty_like(rhs)
for _ in range(2): S.solve_device(ans.data_ptr(), rhs.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
L.fdmb_profile_begin.restype = C.c_int
L.fdmb_profile_end.argtypes = [C.c_char_p, C.c_int]
capi.check(L.fdmb_profile_begin(), "b")
for _ in range(reps): S.solve_device(ans.data_ptr(), rhs.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
buf = C.create_string_buffer(1 << 16)
capi.check(L.fdmb_profile_end(buf, len(buf)), "e")
pts = nx * ny * nz
for line in buf.value.decode().splitlines():
    k, cnt, tot = line.split()
    ms = float(tot) / int(cnt)
    print(f"{k:24s} {ms*1e3:9.1f} us  {16*pts/ms/1e6:7.0f} GB/s")
