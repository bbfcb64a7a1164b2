#!/usr/bin/env python
"""C3 / C4 terrain: the reference's Erosion/lena_gray.png (512 x 512, 8-bit gray, the layout
Grid::LoadHeightfield copies, Erosion/grid.h:98-102: H(x, z) = map[512 * x + z], x = image row) mirrored 2 x 2 to
1024 x 1024 -- exact bytes, no resampling.  The PNG does not travel to the GPU box, so its 512 x 512 bytes are
committed as a fixture and bench.py mirrors them (mirror_2x2 below is the definition both use):
    python tests/golden/make_heightmap.py   ->   tests/golden/lena_gray_512.npz"""
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
img = np.asarray(Image.open("/root/reference/Erosion/lena_gray.png"))
assert img.dtype == np.uint8 and img.shape == (512, 512)      # byte-identical to stbi_load(..., 1) for an 8-bit gray PNG


def mirror_2x2(a):
    top = np.concatenate([a, a[:, ::-1]], axis=1)             # mirror along z: seamless at the joint
    return np.concatenate([top, top[::-1, :]], axis=0)        # mirror along x


big = mirror_2x2(img)
assert big.shape == (1024, 1024) and np.array_equal(big[:512, :512], img)
np.savez_compressed(os.path.join(HERE, "lena_gray_512.npz"), heights_u8=img,
                    source="Erosion/lena_gray.png (min %d, max %d, mean %.2f)" % (img.min(), img.max(), img.mean()))
print("wrote lena_gray_512.npz", img.shape, img.min(), img.max(), round(float(img.mean()), 2))
