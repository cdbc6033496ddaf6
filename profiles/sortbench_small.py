import ctypes as C, sys
sys.path.insert(0,'/root/repo')
import carpedeam_b200 as cd
L=cd.lib()
L.cdg_test_sort_bench.argtypes=[C.c_void_p,C.c_uint64,C.c_int,C.c_int,C.c_int,C.c_int,C.c_int,C.POINTER(C.c_float),C.POINTER(C.c_int)]
ctx=cd.Context(0)
ms=C.c_float(); ps=C.c_int()
print(L.cdg_test_sort_bench(ctx._h,60_000_000,0,16,0,2,1,C.byref(ms),C.byref(ps)), ms.value)
