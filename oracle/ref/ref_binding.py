"""ctypes binding of oracle/_ref/libmon_ref.so (the reference's vendored tiny-cuda-nn + oracle/ref/ref_harness.cu).
TEST / BASELINE INFRASTRUCTURE ONLY: used by oracle/ref/make_golden.py, tests and bench.py --impl reference."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
LIB_PATH = ROOT / "oracle" / "_ref" / "libmon_ref.so"

BASE_JSON = """{
 "loss": {"otype": "Huber"},
 "optimizer": {"otype": "Ema", "decay": 0.95, "nested": {"otype": "ExponentialDecay", "decay_start": 20000, "decay_interval": 10000,
   "decay_base": 0.33, "nested": {"otype": "Adam", "learning_rate": 1e-2, "beta1": 0.9, "beta2": 0.99, "epsilon": 1e-15, "l2_reg": 1e-6}}},
 "encoding": {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 16, "base_resolution": 16},
 "network": {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": %d}
}"""


class RefLib:
    def __init__(self):
        p = LIB_PATH
        if not p.exists():
            raise FileNotFoundError(f"{p}: build it with `make -C oracle/ref` where /root/reference exists")
        L = C.CDLL(str(p))
        vp = C.c_void_p
        L.ref_create.restype = vp
        L.ref_create.argtypes = [C.c_char_p, C.c_uint32]
        L.ref_destroy.argtypes = [vp]
        L.ref_last_error.restype = C.c_char_p
        L.ref_last_error.argtypes = [vp]
        L.ref_n_params.restype = C.c_uint32
        L.ref_n_params.argtypes = [vp]
        L.ref_get.argtypes = [vp, C.c_int, vp]
        L.ref_set_params.argtypes = [vp, vp]
        L.ref_encode.argtypes = [vp, vp, C.c_uint32, vp]
        L.ref_forward.argtypes = [vp, vp, C.c_uint32, vp]
        L.ref_backward.argtypes = [vp, vp, C.c_uint32]
        L.ref_optimizer_step.argtypes = [vp, C.c_float]
        L.ref_inference.argtypes = [vp, vp, C.c_uint32, vp]
        L.ref_scene.argtypes = [vp, C.c_uint32, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, C.c_uint32, vp, vp, vp, C.c_uint8, C.c_int, C.c_uint32]
        L.ref_train.argtypes = [vp, C.c_uint32, vp, vp, vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_uint32)]
        L.ref_last.argtypes = [vp, C.c_int, vp]
        L.ref_is_genuine.restype = C.c_int
        if L.ref_is_genuine():
            L.ref_render.argtypes = [vp] * 12
            L.ref_render2.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_float)]
            L.ref_generate_toc.argtypes = [C.c_float, C.c_float, C.c_float, vp]
            L.ref_marching_cubes.argtypes = [vp, vp, C.c_uint32, C.c_float, vp, vp, C.c_uint32, C.POINTER(C.c_uint32), vp, C.c_uint32, C.POINTER(C.c_uint32)]
        L.ref_use_device.argtypes = [C.c_int]
        self.L = L


class RefModel:
    def __init__(self, n_hidden: int = 1, seed: int = 1337, lib: RefLib | None = None, device: int | None = None):
        self.lib = (lib or RefLib()).L
        if device is not None and self.lib.ref_use_device(device) != 0:   # this host thread drives GPU `device` from here on
            raise RuntimeError(f"cudaSetDevice({device}) failed")
        self.h = self.lib.ref_create((BASE_JSON % n_hidden).encode(), seed)
        if not self.h:
            raise RuntimeError("ref_create failed (see stderr)")
        self.P = int(self.lib.ref_n_params(self.h))

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.ref_last_error(self.h).decode())

    def get(self, which: int) -> np.ndarray:
        out = np.zeros(self.P, np.float32)
        self._ck(self.lib.ref_get(self.h, which, out.ctypes.data))
        return out

    def set_params(self, master: np.ndarray):
        m = np.ascontiguousarray(master, np.float32)
        assert m.size == self.P
        self._ck(self.lib.ref_set_params(self.h, m.ctypes.data))

    def encode(self, pts: np.ndarray) -> np.ndarray:
        pts = np.ascontiguousarray(pts, np.float32)
        out = np.zeros((pts.shape[0], 32), np.uint16)
        self._ck(self.lib.ref_encode(self.h, pts.ctypes.data, pts.shape[0], out.ctypes.data))
        return out

    def forward(self, pts: np.ndarray) -> np.ndarray:
        pts = np.ascontiguousarray(pts, np.float32)
        out = np.zeros((pts.shape[0], 16), np.uint16)
        self._ck(self.lib.ref_forward(self.h, pts.ctypes.data, pts.shape[0], out.ctypes.data))
        return out

    def backward(self, dout_bits: np.ndarray):
        d = np.ascontiguousarray(dout_bits, np.uint16)
        self._ck(self.lib.ref_backward(self.h, d.ctypes.data, d.shape[0]))

    def optimizer_step(self, loss_scale: float = 128.0):
        self._ck(self.lib.ref_optimizer_step(self.h, loss_scale))

    def inference(self, pts: np.ndarray) -> np.ndarray:
        pts = np.ascontiguousarray(pts, np.float32)
        out = np.zeros((pts.shape[0], 4), np.float32)
        self._ck(self.lib.ref_inference(self.h, pts.ctypes.data, pts.shape[0], out.ctypes.data))
        return out


    # ---- whole-iteration loop in the reference's shape (ref_harness.cu: ref_scene / ref_train)
    def scene(self, rgb_list, inst_list, depth_list, poses, H, W, K, boxes_rows, Tow, bmin, bmax, instance_id, use_depth, rays_per_batch):
        n = len(rgb_list)
        self._keep = ([np.ascontiguousarray(a, np.uint8) for a in rgb_list], [np.ascontiguousarray(a, np.uint8) for a in inst_list],
                      [np.ascontiguousarray(a, np.float32) for a in depth_list])
        vp = C.c_void_p
        rgb_p = (vp * n)(*[a.ctypes.data for a in self._keep[0]])
        inst_p = (vp * n)(*[a.ctypes.data for a in self._keep[1]])
        dep_p = (vp * n)(*[a.ctypes.data for a in self._keep[2]])
        poses16 = np.concatenate([np.ascontiguousarray(np.asarray(p, np.float32).reshape(4, 4).T).reshape(16) for p in poses]).astype(np.float32)
        rows = np.ascontiguousarray(np.array(list(boxes_rows), dtype=np.uint32).reshape(-1, 5))
        Kf = np.ascontiguousarray(K, np.float32)
        tow = np.ascontiguousarray(np.asarray(Tow, np.float32).reshape(4, 4).T).reshape(16)
        lo, hi = np.ascontiguousarray(bmin, np.float32), np.ascontiguousarray(bmax, np.float32)
        self.R = rays_per_batch
        self._ck(self.lib.ref_scene(self.h, n, rgb_p, inst_p, dep_p, poses16.ctypes.data, H, W, Kf.ctypes.data, rows.ctypes.data, rows.shape[0],
                                    tow.ctypes.data, lo.ctypes.data, hi.ctypes.data, instance_id, int(use_depth), rays_per_batch))

    def train(self, iters: int, inject=None):
        """Returns (device_ms, wall_ms, loss, n_in) for `iters` iterations of Train_Step's loop body."""
        dm, wm, loss, n_in = C.c_float(0), C.c_float(0), C.c_float(0), C.c_uint32(0)
        if inject is None:
            xy = col = dt = None
        else:
            keep = [np.ascontiguousarray(a, np.float32) for a in inject]
            xy, col, dt = [a.ctypes.data for a in keep]
        self._ck(self.lib.ref_train(self.h, iters, xy, col, dt, C.byref(dm), C.byref(wm), C.byref(loss), C.byref(n_in)))
        return dm.value, wm.value, loss.value, n_in.value

    def last(self, which: int, n: int) -> np.ndarray:
        out = np.zeros(n, np.float32)
        self._ck(self.lib.ref_last(self.h, which, out.ctypes.data))
        return out

    def is_genuine(self) -> bool:
        """True when RO-MAP's glue kernels in the library are the reference's own nerf_model.cu (not the restatement)."""
        return bool(self.lib.ref_is_genuine())

    def render(self, box, Twc, rand_dt):
        """NeRF_Model::Render's device work on box = (FrameId, x, y, h, w) with 64 samples per ray and the EMA weights.
        Returns a dict with the final pixels and every intermediate buffer."""
        b = np.ascontiguousarray(np.array([int(v) for v in box], dtype=np.uint32))
        n, S2 = int(b[3]) * int(b[4]), 64
        twc = np.ascontiguousarray(np.asarray(Twc, np.float32).reshape(4, 4).T).reshape(16)
        dt = np.ascontiguousarray(rand_dt, np.float32)
        assert dt.size == n * S2
        o = dict(rgb=np.zeros((n, 3), np.float32), depth=np.zeros(n, np.float32), mask=np.zeros(n, np.float32), rays=np.zeros((n, 9), np.float32),
                 in_box=np.zeros(n, np.int32), points=np.zeros((n * S2, 3), np.float32), dist=np.zeros(n * S2, np.float32), out4=np.zeros((n * S2, 4), np.float32))
        self._ck(self.lib.ref_render(self.h, b.ctypes.data, twc.ctypes.data, dt.ctypes.data, o["rgb"].ctypes.data, o["depth"].ctypes.data, o["mask"].ctypes.data,
                                     o["rays"].ctypes.data, o["in_box"].ctypes.data, o["points"].ctypes.data, o["dist"].ctypes.data, o["out4"].ctypes.data))
        return o

    def render2(self, box, T, object_centric: bool = False, rand_dt=None, want_rays: bool = False):
        """NeRF_Model::Render (T = camera->world) or one view of RenderVideo (object_centric: T = camera->object, e.g. generate_toc),
        device work + the three D2H copies only; rand_dt None -> cuRAND on the device like the reference.
        Returns dict(rgb, depth, mask, device_ms[, rays, in_box])."""
        b = np.ascontiguousarray(np.array([int(v) for v in box], dtype=np.uint32))
        n = int(b[3]) * int(b[4])
        t16 = np.ascontiguousarray(np.asarray(T, np.float32).reshape(4, 4).T).reshape(16)
        dt = None if rand_dt is None else np.ascontiguousarray(rand_dt, np.float32)
        assert dt is None or dt.size == n * 64
        o = dict(rgb=np.zeros((n, 3), np.float32), depth=np.zeros(n, np.float32), mask=np.zeros(n, np.float32))
        if want_rays:
            o["rays"], o["in_box"] = np.zeros((n, 9), np.float32), np.zeros(n, np.int32)
        ms = C.c_float(0)
        self._ck(self.lib.ref_render2(self.h, b.ctypes.data, t16.ctypes.data, int(object_centric), None if dt is None else dt.ctypes.data,
                                      o["rgb"].ctypes.data, o["depth"].ctypes.data, o["mask"].ctypes.data,
                                      o["rays"].ctypes.data if want_rays else None, o["in_box"].ctypes.data if want_rays else None, C.byref(ms)))
        o["device_ms"] = ms.value
        return o

    def generate_toc(self, theta: float, phi: float, radius: float) -> np.ndarray:
        """NeRF_Model::GenerateToc (the reference's own host code): 4x4 camera -> object pose, row-major numpy."""
        t = np.zeros(16, np.float32)
        self.lib.ref_generate_toc(theta, phi, radius, t.ctypes.data)
        return t.reshape(4, 4).T.copy()

    def marching_cubes(self, density, thresh: float = 2.0):
        """The reference's MarchingCubes + compute_mesh_1ring on a [res,res,res] lattice (z, y, x order = x fastest) over the scene's
        object box.  Returns (verts [n,3] incl. the reference's zero padding to a multiple of 128, normals [n,3] un-normalised, indices [m])."""
        d = np.ascontiguousarray(density, np.float32)
        res = d.shape[0]
        assert d.shape == (res, res, res)
        cap_v, cap_i = 3 * res ** 3, 15 * res ** 3
        verts, normals, idx = np.zeros((cap_v, 3), np.float32), np.zeros((cap_v, 3), np.float32), np.zeros(cap_i, np.uint32)
        nv, ni = C.c_uint32(0), C.c_uint32(0)
        self._ck(self.lib.ref_marching_cubes(self.h, d.ctypes.data, res, thresh, verts.ctypes.data, normals.ctypes.data, cap_v, C.byref(nv),
                                             idx.ctypes.data, cap_i, C.byref(ni)))
        return verts[: nv.value].copy(), normals[: nv.value].copy(), idx[: ni.value].copy()

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None


