#!/usr/bin/env python
"""GPU box: max abs error of __sinf/__cosf per magnitude band (validates e_m in screen_eps)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.binding import RefGpu
ref = RefGpu()
u = 2.0 ** -24
lo = 2.0 ** -20
for hi in (0.5, 1.0, 2.0, 3.1415927, 4.0, 6.2831855, 8.0, 16.0, 32.0, 64.0, 128.0, 1024.0):
    es, ec = ref.fast_trig_err(lo, hi)
    bound = 2.0 ** -20 + 4 * u * hi
    print(f"|x| in [{lo:.3g}, {hi:.6g}]: max err sin {es:.3e} cos {ec:.3e}   budget e_m(hi) = {bound:.3e}  ratio {max(es, ec) / bound:.3f}")
    lo = hi
