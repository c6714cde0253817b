import sys, time, ctypes as C, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import __graft_entry__ as g
import test_oracle_marg as tm
pkg=g.load_package(); abi,synth=pkg.abi,pkg.synth
ctx=pkg.lib.Context(0)
w=synth.make_window(seed=300,K=11,L=150)
for i in range(6):
    t=time.perf_counter(); p=tm.run_marg(abi, ctx.L.bvio_marginalize, w, 0, ctx=ctx.h); print("marg ms", (time.perf_counter()-t)*1e3, p["n"])
