timeout 600 python -m pytest tests/test_lapl_cube_large_gpu.py -m gpu -q -x -k "test_fft_batch_pipe_vs_oracle and 1024" 2>&1 | grep -E "assert|Error|error|^E" | head -20
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
import fdm_b200
from oracle import fdm_oracle as O
N = 1024
for batch in (1, 2, 7, 8, 9, 16, 37):
    x = np.random.default_rng(batch).uniform(-1, 1, (batch, N - 1))
    got = fdm_b200.fft_batch("sFFT", N, x, 0.37, impl="pipe"); want = O.sFFT(x, 0.37)
    err = [O.rel_l2(got[r], want[r]) for r in range(batch)]
    print(batch, ["%.1e" % e for e in err])
PY
