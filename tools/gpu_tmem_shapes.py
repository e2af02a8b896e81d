"""Which (TMEM lane, column) does every register of the tcgen05.ld shapes 16x256b / 16x128b / 16x64b receive?  (fm_debug_tmem_shapes)"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flowmol_b200 import _lib
lib = _lib.load()
out = np.zeros((128, 16), np.int32)
rc = lib.fm_debug_tmem_shapes(out.ctypes.data, 0)
assert rc == 0, lib.fm_last_error()
for name, sl in (("16x256b.x2", slice(0, 8)), ("16x128b.x2", slice(8, 12)), ("16x64b.x2", slice(12, 14))):
    print(name, "(thread: [(lane, column) per register])")
    for t in (0, 1, 2, 3, 4, 5, 8, 16, 31, 32, 33):
        print(f"  thread {t:3d}:", [(int(v) >> 8, int(v) & 255) for v in out[t, sl]])
