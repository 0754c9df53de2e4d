import sys, time, ctypes as C
sys.path.insert(0, ".")
import numpy as np
from colli2de_b200 import scenes, capi
from tests import harness as H
n=100000
b = H.Backend("b200", 1, 120)
sc = scenes.repo_perf_scene(n)
scenes.populate(b, sc)
b.move_translate(sc["ent_id"], sc["first_move"])
b.get_colliding_pairs()
ctx=b.native_context()
lib=capi.lib()
lib.c2d_gpu_query_shapes.argtypes=[C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
shapes=np.arange(n,dtype=np.uint32)
capi.set_profiling(ctx, True)
for rep in range(3):
    rec=C.c_void_p(); cnt=C.c_uint64()
    t=time.perf_counter()
    rc=lib.c2d_gpu_query_shapes(ctx, n, shapes.ctypes.data_as(C.c_void_p), C.byref(rec), C.byref(cnt))
    dt=time.perf_counter()-t
    print("query_shapes all:", rc, cnt.value, f"{dt*1e3:.2f} ms")
print(capi.kernel_profile(ctx))
t=time.perf_counter(); tot,sec=b.get_collisions_many(np.arange(2000,dtype=np.uint32)); print("2000 singles -> table:", tot, f"{sec*1e3:.2f} ms")
t=time.perf_counter(); tot,sec=b.get_collisions_many(np.arange(2000,dtype=np.uint32)); print("2000 lookups:", tot, f"{sec*1e3:.3f} ms")
