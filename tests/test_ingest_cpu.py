"""G-buffer ingest: the oracle's orc_gbuffer_ingest against (a) the torch restatement of the reference's packers
in realism_effects_b200/synth.py on identical inputs and (b) the reference's own packGBuffer / packNormal GLSL
(src/gbuffer/shader/gbuffer_packing.glsl) run through the GLSL runtime."""
import numpy as np
import pytest
import torch

import orc
import refglsl
from realism_effects_b200 import synth


def soa_frame(W=96, H=54, t=1):
    fr = synth.render_frame(W, H, t)
    s = {k: v.cpu().numpy() for k, v in fr.soa.items()}
    return fr, s


def test_oracle_ingest_equals_torch_packers():
    fr, s = soa_frame()
    gb, vel = orc.gbuffer_ingest(s["albedo"], s["normal"], s["material"], s["emissive"], s["motion"], fr.depth.numpy(), normalize_normals=False)
    # the torch packers on the same (quantised) inputs
    diffuse4 = torch.from_numpy(s["albedo"]).float() / 255.0
    mat = torch.from_numpy(s["material"]).float()
    em = torch.from_numpy(s["emissive"]).float()[..., :3]
    nrm = torch.from_numpy(s["normal"])[..., :3]
    want = synth.pack_gbuffer(diffuse4, nrm, mat[..., 0], mat[..., 1], em)
    clear = torch.tensor([0.0, 0.0, 0.0, 1.0])
    bg = fr.background
    want = torch.where(bg.unsqueeze(-1), clear, want).numpy()
    assert gb.view(np.uint32).tobytes() == want.view(np.uint32).tobytes()
    # with fp16-exact material values the ingested planes ARE the generator's planes; the velocity plane always is
    assert vel.view(np.uint32).tobytes() == fr.velocity.numpy().view(np.uint32).tobytes()
    assert (gb[..., :2].view(np.uint32) == fr.gbuffer.numpy()[..., :2].view(np.uint32)).all()      # diffuse + normal words


def test_ingest_options_and_formats():
    fr, s = soa_frame(64, 40)
    d = fr.depth.numpy()
    n16 = s["normal"].astype(np.float16)
    gb_a, vel_a = orc.gbuffer_ingest(s["albedo"], n16, s["material"], None, None, d)                     # no emissive, static, fp16 normals, normalised
    fg = d < 1.0
    assert (vel_a[..., :2] == 0).all() and (gb_a[fg][:, 3].view(np.uint32) == 0).all()
    assert (vel_a[..., 3] == d).all()
    gb_b, vel_b = orc.gbuffer_ingest(s["albedo"].astype(np.float16) / np.float16(255), s["normal"], (s["material"].astype(np.float32) * 255).round().astype(np.uint8),
                                     s["emissive"], s["motion"].astype(np.float16), d, motion_scale=(0.5, 0.5))
    assert np.allclose(vel_b[fg][:, :2], 0.5 * s["motion"][fg][:, :2], rtol=2e-3, atol=1e-7)
    bg = ~fg
    assert (gb_b[bg] == np.array([0, 0, 0, 1], np.float32)).all() and (vel_b[bg] == np.array([0, 0, 0, 1], np.float32)).all()


GLSL_INGEST = """
uniform sampler2D tAlbedo; uniform sampler2D tNormal; uniform sampler2D tMaterial; uniform sampler2D tEmissive;
layout(location = 0) out vec4 oG;
layout(location = 1) out vec4 oN;
%s
void main() {
  vec4 m = textureLod(tMaterial, vUv, 0.);
  vec3 n = textureLod(tNormal, vUv, 0.).xyz;
  oG = packGBuffer(textureLod(tAlbedo, vUv, 0.), n, m.r, m.g, textureLod(tEmissive, vUv, 0.).rgb);
  oN = vec4(packNormal(n), 0., 0., 1.);
}
"""


@pytest.mark.skipif(not refglsl.assemble.available(), reason="needs the reference checkout (the packers are read from it)")
def test_oracle_ingest_equals_the_reference_packgbuffer_glsl():
    fr, s = soa_frame()
    H, W = fr.depth.shape
    glsl = "varying vec2 vUv;\n" + GLSL_INGEST % refglsl.assemble.read("gbuffer/shader/gbuffer_packing.glsl")
    sh = refglsl.Shader("ingest_probe", glsl=glsl)
    sh.tex("tAlbedo", s["albedo"], refglsl.F_RGBA8)
    sh.tex("tNormal", s["normal"], refglsl.F_RGBA32F)
    sh.tex("tMaterial", s["material"], refglsl.F_RGBA16F)
    sh.tex("tEmissive", s["emissive"], refglsl.F_RGBA16F)
    ref_g, ref_n = sh.run(W, H, [(refglsl.F_RGBA32F, None), (refglsl.F_RGBA32F, None)])
    gb, vel = orc.gbuffer_ingest(s["albedo"], s["normal"], s["material"], s["emissive"], s["motion"], fr.depth.numpy(), normalize_normals=False)
    fg = fr.depth.numpy() < 1.0
    assert (gb[fg][:, :3].view(np.uint32) == ref_g[fg][:, :3].view(np.uint32)).all()               # diffuse, normal, roughness/metalness words
    lit = fg & (s["emissive"][..., :3].astype(np.float32).max(-1) > 0)
    assert lit.any() and (gb[lit][:, 3].view(np.uint32) == ref_g[lit][:, 3].view(np.uint32)).all()  # RGBE8 emissive word where the shader's path is defined
    assert (vel[fg][:, 2].view(np.uint32) == ref_n[fg][:, 0].view(np.uint32)).all()                # packNormal
