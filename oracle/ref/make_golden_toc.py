"""Generates tests/golden/romap_toc_golden.npz by RUNNING the reference's own NeRF_Model::GenerateToc
(MON/Core/src/nerf_model.cu:2186-2205, compiled unmodified into oracle/_ref/libmon_ref.so; host code, no GPU needed):
the 60 turn-table poses of RenderVideo (:1834-1846: theta = 6, 12, ... 360 degrees, phi = 30) for three radii.

    LD_LIBRARY_PATH=<dir with libcuda.so.1 (the CUDA stub is enough)> python oracle/ref/make_golden_toc.py

TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
RADII = (0.8, 1.5, 3.25)
PHI = 30.0


def thetas():
    # RenderVideo accumulates in float: cur_theta += 360 / float(60)
    out, cur, step = [], np.float32(0.0), np.float32(360) / np.float32(60)
    for _ in range(60):
        cur = np.float32(cur + step)
        out.append(cur)
    return np.array(out, np.float32)


def main():
    L = C.CDLL(str(ROOT / "oracle" / "_ref" / "libmon_ref.so"))
    L.ref_generate_toc.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p]
    th = thetas()
    toc = np.zeros((len(RADII), len(th), 16), np.float32)
    for i, r in enumerate(RADII):
        for j, t in enumerate(th):
            assert L.ref_generate_toc(float(t), PHI, float(r), toc[i, j].ctypes.data) == 0
    out = ROOT / "tests" / "golden" / "romap_toc_golden.npz"
    np.savez_compressed(out, radii=np.array(RADII, np.float32), phi=np.float32(PHI), thetas=th, toc=toc)
    print("wrote", out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    sys.exit(main())
