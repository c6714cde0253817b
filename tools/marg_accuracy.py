"""How far apart are the device's and the oracle's marginalization (same inputs), in the reference's invariants J^T J and
J^T r -- for both factorizations of the kept information (eigen = the reference's, cholesky = opt-in)?  Development tool."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
sys.path.insert(0, os.path.join(HERE, ".."))
import __graft_entry__ as g  # noqa: E402
import oracle_lib  # noqa: E402
from test_oracle_marg import info_in_state_coords, run_marg  # noqa: E402

pkg = g.load_package()
abi, synth = pkg.abi, pkg.synth
orc = oracle_lib.load()
for mode in ("eigen", "cholesky"):
    if mode == "cholesky":
        os.environ["BVIO_MARG_CHOLESKY"] = "1"
    ctx = pkg.lib.Context(0)
    for seed, K, L in [(0, 11, 150), (3, 11, 400), (5, 11, 150), (6, 11, 1500)]:
        w = synth.make_window(seed=seed, K=K, L=L, prior="frame0")
        pg = run_marg(abi, ctx.L.bvio_marginalize, w, 0, ctx=ctx.h, opts=abi.default_opts())
        po = run_marg(abi, orc.oracle_marginalize, w, 0, opts=abi.default_opts())
        Hg, gg = info_in_state_coords(pg, K, lambda f: f + 1)
        Ho, go = info_in_state_coords(po, K, lambda f: f + 1)
        ev = np.linalg.eigvalsh(Ho)
        print(mode, "seed", seed, "L", L, "n", pg["n"], "dH/|H| %.2e" % (np.abs(Hg - Ho).max() / np.abs(Ho).max()),
              "dg/|g| %.2e" % (np.abs(gg - go).max() / max(np.abs(go).max(), 1.0)), "cond %.1e" % (ev.max() / max(ev.min(), 1e-300)),
              "J rows gpu/oracle", pg["J"].shape[0], po["J"].shape[0])
    ctx.close()
