import sys, time, dataclasses
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import __graft_entry__ as g
import test_oracle_marg as tm
pkg=g.load_package(); abi,synth=pkg.abi,pkg.synth
ctx=pkg.lib.Context(0)
keys=("n","block_kind","block_frame","block_idx","x0","lin_jac","lin_res")
p0=tm.run_marg(abi, ctx.L.bvio_marginalize, synth.make_window(seed=300,K=11,L=150), 0, ctx=ctx.h)
w=dataclasses.replace(synth.make_window(seed=301,K=11,L=150), prior={k:p0[k] for k in keys})
for i in range(5):
    t0=time.perf_counter(); job=abi.MarginalizeJob(ctx.L, ctx.h, w, 0); t1=time.perf_counter(); p=job.end(); t2=time.perf_counter()
    print("begin %.3f ms  end %.3f ms  (inner begin %.3f end %.3f)"%((t1-t0)*1e3,(t2-t1)*1e3, job.t_begin*1e3, job.t_end*1e3))
for i in range(3):
    t0=time.perf_counter(); p=tm.run_marg(abi, ctx.L.bvio_marginalize, w, 0, ctx=ctx.h); print("sync %.3f ms (ffi %.3f)"%((time.perf_counter()-t0)*1e3, abi.call_marginalize.t_call*1e3))
