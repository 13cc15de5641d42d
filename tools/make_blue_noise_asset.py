"""Decode the reference's blue-noise PNG into the raw RGBA8 texel array the engine uploads.

The texture is an *input* of every stochastic pass (src/utils/BlueNoiseUtils.js:6-15); the
reference loads it with three.js' TextureLoader (flipY = true, NoColorSpace), i.e. GL texel row 0
is the LAST image row.  This script reproduces that memory layout once, in the container that has
/root/reference, and writes realism_effects_b200/assets/blue_noise_rgba_128.bin (65536 bytes).
"""
import hashlib
import os
import sys

import numpy as np
from PIL import Image

SRC = "/root/reference/src/utils/blue_noise_rgba.png"
SHA = "359e71ac175818b5ec4b781342235d46b498aedf0b48b45c5be8e0345fe83713"
DST = os.path.join(os.path.dirname(__file__), "..", "realism_effects_b200", "assets", "blue_noise_rgba_128.bin")

raw = open(SRC, "rb").read()
assert hashlib.sha256(raw).hexdigest() == SHA, "unexpected blue-noise PNG"
im = np.asarray(Image.open(SRC).convert("RGBA"), dtype=np.uint8)
assert im.shape == (128, 128, 4)
gl = np.ascontiguousarray(im[::-1])  # flipY
gl.tofile(DST)
print("wrote", os.path.abspath(DST), "sha256", hashlib.sha256(gl.tobytes()).hexdigest(), "mean", gl.reshape(-1, 4).mean(0), file=sys.stderr)
