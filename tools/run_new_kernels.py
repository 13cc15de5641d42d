"""Launches the round-2 streaming kernels once each at 3840x2160 (for `ncu -k regex:gbuffer_ingest_kernel|effects_kernel|taa_kernel`):
    ncu --set full --clock-control none --import-source on -k 'regex:gbuffer_ingest|effects_kernel|taa_kernel' -f -o gpurun_out/r02_new_kernels python tools/run_new_kernels.py"""
import os
import sys

import numpy as np
import torch  # noqa: F401  (initialises CUDA before the engine)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import chain_harness as ch  # noqa: E402
from realism_effects_b200 import abi, engine, synth  # noqa: E402

W, H = 3840, 2160
fr = synth.render_frame(W, H, 1, device="cuda:0")
ctx = engine.Context(0, synth.load_blue_noise())
s = {k: v.cpu().numpy() for k, v in fr.soa.items()}
d = ctx.upload(fr.depth.cpu().numpy())
vel = ctx.upload(fr.velocity.cpu().numpy())
dl = ctx.upload(fr.direct_light.cpu().numpy())
planes = [ctx.upload(s["albedo"]), ctx.upload(s["normal"].astype(np.float16)), ctx.upload((s["material"].astype(np.float32) * 255).round().astype(np.uint8)),
          ctx.upload(s["emissive"]), ctx.upload(s["motion"].astype(np.float16))]
og, ov = ctx.alloc(abi.FMT_RGBA32F, W, H), ctx.alloc(abi.FMT_RGBA32F, W, H)
fxo = ctx.alloc(abi.FMT_RGBA16F, W, H)
th, to = ctx.alloc(abi.FMT_RGBA8, W, H), ctx.alloc(abi.FMT_RGBA8, W, H)
fxp = ch.fx_params(fr.cam.uniforms(), [abi.FX_SHARPNESS, abi.FX_GRADUAL_BACKGROUND, abi.FX_SPARKLE])
tp = abi.TaaParams()
tp.camera_not_moved_frames, tp.srgb_output = 3.0, 1
for _ in range(3):
    ctx.gbuffer_ingest(*planes, d, og, ov)
    ctx.effects(fxp, dl, d, vel, fxo)
    ctx.taa(tp, dl, th, to)
ctx.sync()
print("done")
