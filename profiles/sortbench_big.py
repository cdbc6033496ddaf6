"""sorts beyond 2^30 records (64-bit tile states): order, stability and checksums are verified inside cdg_test_sort_bench"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carpedeam_b200 as cd
L = cd.lib()
L.cdg_test_sort_bench.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]
ctx = cd.Context(0)
for n, has_val, bits, variant in ((290_000_000, 0, 56, 8), (345_000_000, 1, 40, 0), (1_300_000_000, 0, 56, 8), (1_200_000_000, 1, 40, 0), (3_000_000_000, 0, 16, 8)):
    ms = C.c_float(); ps = C.c_int()
    rc = L.cdg_test_sort_bench(ctx._h, n, 0, bits, has_val, variant, 1, C.byref(ms), C.byref(ps))
    rec = 16 if has_val else 8
    gb = n * rec * (2 * ps.value + 1) / 1e9
    msg = "" if rc == 0 else L.cdg_last_error(ctx._h).decode()
    print("n=%d has_val=%d variant=%d rc=%d passes=%d ms=%.2f  %.0f GB/s (%.1f%% of 6537) %s" % (n, has_val, variant, rc, ps.value, ms.value, gb / max(ms.value, 1e-9) * 1e3, gb / max(ms.value, 1e-9) * 1e3 / 65.376, msg), flush=True)
