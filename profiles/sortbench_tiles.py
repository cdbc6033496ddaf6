"""sort micro-benchmark over tile shapes (cdg_test_sort_bench): variant 0 = 256 threads x 16 keys, 2 = 512 x 8, 8 = 512 x 16"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carpedeam_b200 as cd
L = cd.lib()
L.cdg_test_sort_bench.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]
ctx = cd.Context(0)
n = 290_000_000
for has_val, bits in ((0, 56), (1, 40)):
    for variant in (0, 2, 8):
        ms = C.c_float(); ps = C.c_int()
        rc = L.cdg_test_sort_bench(ctx._h, n, 0, bits, has_val, variant, 3, C.byref(ms), C.byref(ps))
        rec = 16 if has_val else 8
        gb = n * rec * (2 * ps.value + 1) / 1e9
        print("has_val=%d variant=%d rc=%d passes=%d ms=%.2f  %.0f GB/s (%.1f%% of 6537)" % (has_val, variant, rc, ps.value, ms.value, gb / ms.value * 1e3, gb / ms.value * 1e3 / 65.376))
