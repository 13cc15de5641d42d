"""Convert the reference demo's environment map (example/public/hdr/spree_bank_1k.hdr, the one example/main.js:278 loads; a Poly Haven CC0
panorama) into realism_effects_b200/assets/spree_bank_1k_rgbe.npz: the raw RGBE8 texels (H, W, 4) uint8 in FILE order (top scanline first).
SURVEY.md §8(d) names this map as the bench environment.  Run in the build container:  python tools/make_env_asset.py"""
import os

import numpy as np

SRC = os.path.join(os.environ.get("RFX_REFERENCE_DIR", "/root/reference"), "example", "public", "hdr", "spree_bank_1k.hdr")
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "realism_effects_b200", "assets", "spree_bank_1k_rgbe.npz")


def read_rgbe(path):
    """Radiance .hdr (new-style RLE scanlines) -> (H, W, 4) uint8 RGBE"""
    with open(path, "rb") as f:
        buf = f.read()
    pos = buf.index(b"\n\n") + 2
    head = buf[:pos].decode("ascii", "replace")
    assert "#?RADIANCE" in head or "#?RGBE" in head, head[:40]
    end = buf.index(b"\n", pos)
    res = buf[pos:end].decode().split()
    assert res[0] == "-Y" and res[2] == "+X", res
    H, W = int(res[1]), int(res[3])
    pos = end + 1
    out = np.empty((H, W, 4), np.uint8)
    for y in range(H):
        assert buf[pos] == 2 and buf[pos + 1] == 2 and ((buf[pos + 2] << 8) | buf[pos + 3]) == W, "not a new-style RLE scanline"
        pos += 4
        for c in range(4):
            x = 0
            while x < W:
                n = buf[pos]
                pos += 1
                if n > 128:
                    n -= 128
                    out[y, x:x + n, c] = buf[pos]
                    pos += 1
                else:
                    out[y, x:x + n, c] = np.frombuffer(buf, np.uint8, n, pos)
                    pos += n
                x += n
    return out


if __name__ == "__main__":
    rgbe = read_rgbe(SRC)
    np.savez_compressed(DST, rgbe=rgbe)
    print("wrote", DST, os.path.getsize(DST), "bytes", rgbe.shape)
