import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from smore_b200 import capi
L = capi.lib()
capi.check(L.smore_init(0))
t, p, s = C.c_int(), C.c_int(), C.c_int()
capi.check(L.smore_debug_sm_partition(int(sys.argv[1]), C.byref(t), C.byref(p), C.byref(s)))
print(f"reserve {sys.argv[1]}: device {t.value} SMs, partition {p.value}, probe kernel ran on {s.value} distinct SMs")
