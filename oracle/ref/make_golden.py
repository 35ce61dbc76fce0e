"""Generates tests/golden/tcnn_golden.npz by RUNNING the reference's own vendored tiny-cuda-nn
(oracle/_ref/libmon_ref.so, built from /root/reference by oracle/ref/Makefile) on a GPU.

    python oracle/ref/make_golden.py [out.npz]          # needs a CUDA device; run under gpurun

Inputs are deterministic (pcg32 streams / numpy default_rng with fixed seeds) so the CPU tests can
regenerate them bit-for-bit and compare the oracle (oracle/mon_oracle.cpp) with what the reference computed:
rows A4 (hash encode), A5 (MLP forward), A8 (MLP backward), A9 (grid scatter), A10/A11 (Adam + EMA) and
A12 (parameter initialisation) of SURVEY.md §8a.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))

from ref_binding import BASE_JSON, RefLib, RefModel  # noqa: E402,F401


# ---- deterministic inputs shared with tests/test_golden_tcnn.py ---------------------------------
def golden_points(n: int = 1024) -> np.ndarray:
    rng = np.random.default_rng(2024)
    pts = rng.random((n, 3), dtype=np.float32)
    pts[:8] = [[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0.99999, 0.5, 1e-7], [0.25, 0.75, 0.125]]
    return pts


def golden_master(P: int, n_mlp: int) -> np.ndarray:
    """Parameters with O(1) table entries so that every level contributes visibly (init values are +-1e-4)."""
    rng = np.random.default_rng(7)
    m = rng.uniform(-1.0, 1.0, P).astype(np.float32)
    m[:n_mlp] *= 0.3
    return m


def golden_dout(n: int) -> np.ndarray:
    rng = np.random.default_rng(99)
    d = np.zeros((n, 16), np.float32)
    d[:, :4] = rng.normal(0, 2e-2, (n, 4)).astype(np.float32)
    d[n // 2:, :] *= (rng.random((n - n // 2, 1)) > 0.3)  # rows of exact zeros, like samples after the early stop
    return d.astype(np.float16).view(np.uint16)


SAMPLE_STRIDE = 997  # grid parameters are sampled with this stride to keep the fixture small


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main(out_path: str):
    lib = RefLib()
    gold = {}
    for n_hidden in (1, 2):
        tag = f"h{n_hidden}_"
        m = RefModel(n_hidden, 1337, lib)
        n_mlp = 64 * 32 + (n_hidden - 1) * 64 * 64 + 16 * 64
        P = m.P
        # A12: initialisation
        init = m.get(0)
        gold[tag + "init_mlp"] = init[:n_mlp].copy()
        gold[tag + "init_grid_sample"] = init[n_mlp::SAMPLE_STRIDE].copy()
        gold[tag + "init_sha256"] = np.array(sha(init))
        gold[tag + "init_fp16_sample"] = m.get(1)[n_mlp::SAMPLE_STRIDE].copy()
        # A4 / A5 on O(1) parameters
        master = golden_master(P, n_mlp)
        m.set_params(master)
        pts = golden_points(1024)
        gold[tag + "enc"] = m.encode(pts)
        gold[tag + "out"] = m.forward(pts)
        # A8 / A9
        dout = golden_dout(1024)
        m.backward(dout)
        grad = m.get(3)
        gold[tag + "grad_mlp"] = grad[:n_mlp].copy()
        nz = np.flatnonzero(grad[n_mlp:])
        gold[tag + "grad_grid_nnz"] = np.array(nz.size)
        gold[tag + "grad_grid_idx"] = nz[::37].astype(np.uint32)
        gold[tag + "grad_grid_val"] = grad[n_mlp:][nz[::37]].copy()
        gold[tag + "grad_grid_abs_sum"] = np.array(np.abs(grad[n_mlp:].astype(np.float64)).sum())
        # A10 / A11: three optimizer steps on the same gradient
        watch = np.concatenate([np.arange(n_mlp), n_mlp + nz[::37], n_mlp + np.arange(0, P - n_mlp, SAMPLE_STRIDE * 8)]).astype(np.int64)
        gold[tag + "watch"] = watch
        for s in (1, 2, 3):
            m.optimizer_step(128.0)
            gold[tag + f"step{s}_master"] = m.get(0)[watch].copy()
            gold[tag + f"step{s}_fp16"] = m.get(1)[watch].copy()
            gold[tag + f"step{s}_ema"] = m.get(2)[watch].copy()
        # inference with the EMA weights (what Render uses)
        gold[tag + "infer"] = m.inference(pts)
        m.close()
    np.savez_compressed(out_path, **gold)
    print("wrote", out_path, {k: (v.shape, str(v.dtype)) for k, v in gold.items() if k.startswith("h1_")})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "gpurun_out" / "tcnn_golden.npz"))
