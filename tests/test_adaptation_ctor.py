"""AdaptationUnifRW constructor semantics -- the reference's testset
"adaptation for random walk" (test/runtests.jl:34-85) line by line."""
import numpy as np

import extensiblemcmc_jl_b200 as em
from extensiblemcmc_jl_b200 import _abi


def _template():
    t = em.AdaptationUnifRW(1.0)
    assert (t.target_accpt_rate, t.adapt_every_k_steps, t.scale, t.min, t.max, t.offset, t.N) == (
        0.234, 100, 1.0, 1.0e-12, 1e7, 1e2, 1)           # runtests.jl:35
    assert (t.proposed, t.accepted) == (0, 0)
    return t


def test_defaults_and_scalar_promotion():
    template = _template()
    assert template == em.AdaptationUnifRW(1.0)
    assert template == em.AdaptationUnifRW([2.0])
    assert template == em.AdaptationUnifRW(np.array([3.0]))

    longer = em.AdaptationUnifRW([1.0, 2.0])
    assert template != longer
    assert em.isequal_except(template, longer, "N")
    longer3 = em.AdaptationUnifRW(np.array([1.0, 2.0, 3.0]))
    assert template != longer3
    assert em.isequal_except(template, longer3, "N")

    new_scale = em.AdaptationUnifRW(1.0, scale=3.0)
    assert template != new_scale
    assert em.isequal_except(template, new_scale, "scale")

    new_params = em.AdaptationUnifRW(1.0, scale=3.0, target_accpt_rate=0.111, min=10.0)
    assert em.isequal_except(template, new_params, "scale", "target_accpt_rate", "min")
    assert new_params.scale == 3.0
    assert new_params.target_accpt_rate == 0.111
    assert new_params.min == 10.0


def test_vector_promotion():
    ar_vec = em.AdaptationUnifRW([1.0, 2.0], scale=[3.0, 4.0], target_accpt_rate=0.111, min=10.0)
    assert ar_vec.target_accpt_rate == 0.111
    assert np.array_equal(ar_vec.min, [10.0, 10.0])
    assert np.array_equal(ar_vec.max, [1e7, 1e7])
    assert np.array_equal(ar_vec.scale, [3.0, 4.0])
    assert np.array_equal(ar_vec.offset, [100.0, 100.0])
    assert ar_vec.N == 2
    assert ar_vec.adapt_every_k_steps == 100
    ar_vec2 = em.AdaptationUnifRW([1.0, 2.0], scale=[3.0, 4.0], target_accpt_rate=0.111, min=[10.0, 10.0])
    assert ar_vec == ar_vec2


def test_abi_translation_and_vector_rejection():
    a = em.AdaptationUnifRW([0.0], adapt_every_k_steps=50, scale=0.1).to_abi()
    assert (a.kind, a.adapt_every_k_steps, a.target_accpt_rate, a.scale, a.min, a.max, a.offset) == (
        _abi.ADAPT_UNIF_RW, 50, 0.234, 0.1, 1e-12, 1e7, 1e2)
    import pytest
    with pytest.raises(NotImplementedError):
        # vector scale is constructible but cannot run in the reference either (adaptation.jl:312-319)
        em.AdaptationUnifRW([1.0, 2.0], scale=[3.0, 4.0]).to_abi()
    assert em.NoAdaptation().to_abi().kind == _abi.ADAPT_NONE
