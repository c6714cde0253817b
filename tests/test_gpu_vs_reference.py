"""GPU parity against the REFERENCE's own code, directly: libbvio.so next to oracle/_ref/libvins_ref.so (the reference's
sources compiled here from /root/reference, see tests/test_reference_pin.py; the prebuilt library travels to the GPU
box).  Skipped when that library is absent."""
import ctypes as C

import numpy as np
import pytest

import ref_lib
from test_oracle_marg import info_in_state_coords, run_marg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    L = ref_lib.load()
    if L is None:
        pytest.skip("oracle/_ref/libvins_ref.so not built and /root/reference not present")
    return L


@pytest.mark.parametrize("seed,N,U,n_lm,kappa", [(0, 120, 0, 60, 25), (1, 150, 12, 80, 30), (3, 200, 20, 120, 40)])
def test_cuda_select_matches_reference_select(pkg, ref, seed, N, U, n_lm, kappa):
    """bvio_select vs FeatureSelector::select (the reference's code, nanoflann included): same ids, same order."""
    from test_reference_pin import reference_select_case
    abi = pkg.abi
    ref_ids, prob = reference_select_case(pkg, ref, seed, N, U, n_lm, kappa)
    ctx = pkg.lib.Context(0)
    hs, ss = abi.SelectHandle(prob), abi.SelectSummary()
    out = np.zeros(kappa, np.int32)
    ctx.check(ctx.L.bvio_select(ctx.h, C.byref(hs.s), abi.iptr(out), None, C.byref(ss)), "bvio_select")
    ctx.close()
    assert ss.n_selected == len(ref_ids) and (out[:ss.n_selected] == ref_ids).all(), (out[:ss.n_selected], ref_ids)


@pytest.mark.parametrize("seed,K,L", [(0, 11, 150), (1, 6, 40)])
def test_cuda_marginalize_matches_reference_marginalize(pkg, ref, seed, K, L):
    """bvio_marginalize vs MarginalizationInfo::marginalize on the reference's residual blocks: the new prior's
    quadratic form in state coordinates (tolerances of tests/test_gpu_marg.py)."""
    abi, synth = pkg.abi, pkg.synth
    w = synth.make_window(seed=seed, K=K, L=L)
    ctx = pkg.lib.Context(0)
    pg = run_marg(abi, ctx.L.bvio_marginalize, w, 0, ctx=ctx.h)
    ctx.close()
    pr = run_marg(abi, ref.ref_marginalize, w, 0)
    assert pg["n"] == pr["n"]
    Hg, gg = info_in_state_coords(pg, K, lambda f: f + 1)
    Hr, gr = info_in_state_coords(pr, K, lambda f: f + 1)
    assert np.abs(Hg - Hr).max() <= 1e-7 * np.abs(Hr).max()
    assert np.abs(gg - gr).max() <= 5e-5 * max(np.abs(gr).max(), 1.0)
