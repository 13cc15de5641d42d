"""OrthographicCamera through the CUDA chain (the `#else` branches of PERSPECTIVE_CAMERA: getViewZ in K1 / K2 / K3 / K4 — the only places the reference's
shaders branch on the camera type, ssgi_utils.frag:8, reproject.frag:14, denoiser_compose_functions.glsl:4 — plus K1's general, non-sparse projection).

The ORACLE side of this configuration is pinned against the reference's own shaders (tests/test_reference_glsl.py, tools/pin_oracle.py cases
*_orthographic).  This GPU test was written after the round's GPU minutes were spent, so it had not run on hardware when it was committed: it is a
non-strict xfail (XPASS = the CUDA path conforms, xfail = it does not; either way the run goes on), and the file sorts last."""
import pytest

import chain_harness as ch
from realism_effects_b200 import abi

pytestmark = pytest.mark.gpu


@pytest.mark.xfail(strict=False, reason="added without a GPU run (budget spent); see the module docstring")
@pytest.mark.parametrize("fast", [True, False])
@pytest.mark.parametrize("mode", [abi.MODE_SSGI, abi.MODE_SSR])
def test_chain_parity_orthographic_camera(built, mode, fast):
    res = ch.run_chain_parity(width=160, height=90, frames=3, fast_math=fast, inputs_kw=dict(orthographic=True), mode=mode)
    print(res["summary"])
    assert res["ok"], res["summary"]
