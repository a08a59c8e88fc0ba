import sys, time, os
sys.path.insert(0, '/root/repo')
os.environ["DCMA_B200_TRACE"] = "1"
import torch, numpy as np, ctypes as C
from dicomautomaton_b200 import capi, phantom
from dicomautomaton_b200.demons import AlignViaDemonsParams
shape = (256, 512, 512)
f, m = phantom.make_pair(shape, device="cuda")
hf, hm = f.cpu().pin_memory(), m.cpu().pin_memory()
hD = torch.empty((256*512*512, 3), dtype=torch.float64).pin_memory()
par = AlignViaDemonsParams(max_iterations=200, convergence_threshold=0.0, deformation_field_smoothing_sigma=1.5, update_field_smoothing_sigma=1.5, use_diffeomorphic=True, verbosity=0, field_mode="fp64")
geom = capi.Geom(256, 512, 512, 1, 1.0, 1.0, 1.0); cp = par.to_c(); err = C.create_string_buffer(512); st = capi.RunStats(); lib = capi.lib()
for i in range(3):
    t0 = time.perf_counter()
    capi.check(lib.dcma_b200_demons_register(C.byref(geom), hf.data_ptr(), hm.data_ptr(), C.byref(cp), hD.data_ptr(), None, C.byref(st), err, len(err)), err)
    print("call wall %.1f ms" % (1e3 * (time.perf_counter() - t0)), flush=True)
