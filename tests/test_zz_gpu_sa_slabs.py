"""The packed (symmetric/antisymmetric) pp-ladder with the output-pair space split into several column
blocks -- the multi-rank code path of cc.cu (pack from a slab with b0 != 0, block GEMMs into one
[L+|L-] buffer, unpack per slab) -- exercised on ONE GPU through jues_b200_sa_ladder and compared with
the dense contraction.  (The file name sorts last on purpose: these are kernel-level checks of the
sharded layout; the end-to-end multi-GPU parity check is tools/dist_check.py.)"""
import numpy as np
import pytest

import jues.jl_b200 as jb

pytestmark = pytest.mark.gpu


def inputs(o, v, seed):
    rng = np.random.default_rng(seed)
    tau = rng.standard_normal((o, o, v, v))
    W = rng.standard_normal((v, v, v, v))
    W = W + W.transpose(1, 0, 3, 2)               # <ef|ab> = <fe|ba>
    return tau, W


@pytest.mark.parametrize("o,v", [(3, 5), (4, 12), (5, 19), (8, 32), (2, 2)])
@pytest.mark.parametrize("nslabs", [1, 2, 3, 4, 8])
def test_sa_ladder_in_column_blocks(ctx, o, v, nslabs):
    tau, W = inputs(o, v, 100 * o + v)
    ref = np.einsum("ijef,efab->ijab", tau, W, optimize=True)
    got = ctx.sa_ladder(tau, W, nslabs=nslabs)
    assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), (o, v, nslabs)


def test_sa_ladder_matches_the_numpy_mirror(ctx):
    """Same numbers as the numpy statement of the device algorithm (tests/factorized_model.py)."""
    import factorized_model as fm
    tau, W = inputs(3, 8, 5)
    for nr in (1, 2, 4):
        got = ctx.sa_ladder(tau, W, nslabs=nr)
        assert np.abs(got - fm.sa_ladder(tau, W, nranks=nr)).max() <= 1e-12


def test_sa_ladder_argument_errors(ctx):
    tau, W = inputs(3, 4, 1)
    with pytest.raises(jb.JuesError):
        ctx.sa_ladder(tau, W, nslabs=0)
    with pytest.raises(jb.JuesError):
        ctx.sa_ladder(tau, W[:, :, :, :3], nslabs=1)
