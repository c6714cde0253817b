"""How much does Omega_PRIOR from the back end (bvio_window_omega_prior, opt-in) change what the selector picks, compared
with the reference's I9 (feature_selector.cpp:602-609)?  One closed-loop session driven with the reference behaviour; at
every frame the same selection problem is ALSO solved with the window's Omega_PRIOR and the two id lists are compared.
Usage: python tools/omega_prior_effect.py [frames]"""
import ctypes as C
import dataclasses
import json
import sys

import numpy as np

sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as g

pkg = g.load_package()
abi, sl = pkg.abi, pkg.slider
ctx = pkg.lib.Context(0)
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 300


class Both(sl.GpuBackend):
    stats = []

    def select(self, prob):
        ids = super().select(prob)
        w = self.sim.build_window(self)[0]
        om = self.omega_prior(w, self.sim.opts)
        ids2 = super().select(dataclasses.replace(prob, omega_prior=om))
        a, b = set(ids.tolist()), set(ids2.tolist())
        same_prefix = 0
        for x, y in zip(ids.tolist(), ids2.tolist()):
            if x != y:
                break
            same_prefix += 1
        self.stats.append((len(a), len(a & b) / max(len(a | b), 1), same_prefix, float(np.linalg.eigvalsh(0.5 * (om + om.T)).min()),
                           float(np.trace(om))))
        return ids


sim = sl.SlidingWindowSim(seed=7, max_feats=150, max_cand=300, H=10, frame_dt=1.0 / 30.0)
sim.opts, sim.track_loss = dict(strategy=1), 0.2
be = Both(ctx, abi)
be.sim = sim
for f in range(frames + sim.K + 4):
    sim.step(be)
st = np.array(be.stats)
out = {"frames": len(st), "mean_kappa": float(st[:, 0].mean()), "mean_jaccard_overlap": float(st[:, 1].mean()),
       "frames_identical_set": float((st[:, 1] == 1.0).mean()), "mean_identical_prefix": float(st[:, 2].mean()),
       "omega_prior_min_eig_median": float(np.median(st[:, 3])), "omega_prior_trace_median": float(np.median(st[:, 4]))}
print(json.dumps(out))
ctx.close()
