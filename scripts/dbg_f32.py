"""Device-resident timing of the single-precision LaplCube / NSCube next to the fp64 ones."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fdm_b200
L = fdm_b200.lib()
for n in (127, 255, 511):
    dx = 1.0 / n; l = 1 + dx
    for dt, cls in ((torch.float32, fdm_b200.LaplCubeF32), (torch.float64, fdm_b200.LaplCube)):
        S = cls(dx, dx, dx, l, l, l, n, n, n)
        rhs = torch.rand(n ** 3, dtype=dt, device="cuda") - 0.5
        ans = torch.empty_like(rhs)
        st = torch.cuda.Stream(); torch.cuda.synchronize()
        for _ in range(5): S.solve_device(ans.data_ptr(), rhs.data_ptr(), st.cuda_stream)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(50): S.solve_device(ans.data_ptr(), rhs.data_ptr(), st.cuda_stream)
        e1.record(st); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        print(f"LaplCube {n}^3 {str(dt):14s} {ms * 1e3:9.1f} us  {n ** 3 / 1e9 / (ms * 1e-3):7.2f} Gpts/s")
import time
for n in (127, 255):
    for cls in (fdm_b200.NSCubeF32, fdm_b200.NSCube):
        ns = cls(nx=n, nz=n, Re=1000.0, dt=0.005)
        ns.step(5)
        t0 = time.perf_counter(); ns.step(100); dtm = (time.perf_counter() - t0) / 100
        print(f"{cls.__name__:10s} {n}^3  {dtm * 1e3:8.3f} ms/step  {1 / dtm:8.1f} steps/s")
