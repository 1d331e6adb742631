"""Development aid: forward / inverse 3-D FFT against numpy for several sizes, reporting where errors sit."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genetic_b200 as gb

for n in [int(a) for a in sys.argv[1:]] or [256, 512]:
    rng = np.random.default_rng(1)
    x = rng.standard_normal((n, n, n))
    g = gb.Grid(100.0, n)
    ref = np.fft.rfftn(x) / n ** 1.5
    for rep in range(3):
        f = gb.Field(g, x, fourier=False)
        f.toFourier()
        F = f.getDataVector()
        err = np.abs(F - ref)
        bad = np.argwhere(err > 1e-9)
        print("n=%d rep %d forward: max err %.3e, bad slots %d" % (n, rep, err.max(), len(bad)), flush=True)
        if len(bad):
            print("   bad kx range", bad[:, 0].min(), bad[:, 0].max(), "ky", bad[:, 1].min(), bad[:, 1].max(), "kz",
                  bad[:, 2].min(), bad[:, 2].max(), "first", bad[:5].tolist())
        f2 = gb.Field(g, ref.copy(), fourier=True)
        f2.toReal()
        y = f2.getDataVector()
        err = np.abs(y - x)
        bad = np.argwhere(err > 1e-9)
        print("n=%d rep %d inverse: max err %.3e, bad cells %d" % (n, rep, err.max(), len(bad)), flush=True)
        if len(bad):
            print("   bad x range", bad[:, 0].min(), bad[:, 0].max(), "y", bad[:, 1].min(), bad[:, 1].max(), "z",
                  bad[:, 2].min(), bad[:, 2].max(), "first", bad[:5].tolist())
