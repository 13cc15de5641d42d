"""GPU parity of the whole SSGI chain (K1 -> K2 -> K3 x n -> K4) through the C ABI vs the oracle."""
import numpy as np
import pytest

import chain_harness as ch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kw", [
    dict(width=192, height=108, frames=3),
    dict(width=161, height=91, frames=2, denoise_iterations=2),            # odd sizes: helper lanes beyond the edge
    dict(width=128, height=72, frames=2, importance_sampling=False),
    dict(width=128, height=72, frames=2, use_envmap=False, use_direct_light=False, steps=8, refine_steps=0),
    dict(width=128, height=72, frames=2, missed_rays=True),
])
def test_chain_parity(built, kw):
    res = ch.run_chain_parity(**kw)
    print(res["summary"])
    assert res["ok"], res["summary"]
    assert res["launches"] > 0
