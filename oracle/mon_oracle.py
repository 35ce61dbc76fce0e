"""ctypes binding of the CPU ORACLE (oracle/mon_oracle.cpp -> oracle/_build/libmon_oracle.so).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs — never by anything under ro_map_b200/.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_build" / "libmon_oracle.so"


class OrcConfig(C.Structure):
    _fields_ = [
        ("n_levels", C.c_uint32), ("n_features", C.c_uint32), ("log2_hashmap_size", C.c_uint32), ("base_resolution", C.c_uint32),
        ("per_level_scale", C.c_float),
        ("n_neurons", C.c_uint32), ("n_hidden_layers", C.c_uint32), ("padded_output_width", C.c_uint32),
        ("learning_rate", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("epsilon", C.c_float), ("l2_reg", C.c_float),
        ("ema_decay", C.c_float),
        ("decay_start", C.c_uint32), ("decay_interval", C.c_uint32), ("decay_base", C.c_float),
        ("loss_scale", C.c_float),
    ]


class OrcBbox2d(C.Structure):
    _fields_ = [("FrameId", C.c_uint32), ("x", C.c_uint32), ("y", C.c_uint32), ("h", C.c_uint32), ("w", C.c_uint32)]


class OrcRay(C.Structure):
    _fields_ = [("o", C.c_float * 3), ("d", C.c_float * 3), ("d_norm", C.c_float), ("tmin", C.c_float), ("tmax", C.c_float)]


class OrcFrame(C.Structure):
    _fields_ = [("rgb", C.c_void_p), ("instance", C.c_void_p), ("depth", C.c_void_p), ("pose", C.c_float * 16)]


def build(force: bool = False) -> Path:
    src = [HERE / "mon_oracle.cpp", HERE / "mon_oracle.h"]
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < max(p.stat().st_mtime for p in src):
        subprocess.run(["make", "-C", str(HERE), "-s"], check=True)
    return LIB_PATH


_lib = None
_vp = C.c_void_p


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB_PATH))
        L.orc_grid_layout.restype = C.c_uint32
        L.orc_n_mlp_params.restype = C.c_uint32
        L.orc_n_params.restype = C.c_uint32
        L.orc_f2h.restype = C.c_uint16
        L.orc_f2h.argtypes = [C.c_float]
        L.orc_h2f.restype = C.c_float
        L.orc_h2f.argtypes = [C.c_uint16]
        L.orc_ray_intersect.restype = C.c_int
        L.orc_generate_rays.restype = C.c_uint32
        L.orc_object_create.restype = _vp
        L.orc_object_create.argtypes = [C.POINTER(OrcConfig), C.c_uint32, C.c_uint32, C.c_uint32, _vp, _vp, _vp, C.c_uint8, C.c_int, C.c_int]
        L.orc_object_destroy.argtypes = [_vp]
        L.orc_object_n_params.restype = C.c_uint32
        L.orc_object_n_params.argtypes = [_vp]
        L.orc_object_get.argtypes = [_vp, C.c_int, _vp]
        L.orc_object_set_params.argtypes = [_vp, _vp]
        L.orc_object_train_iter.restype = C.c_float
        L.orc_object_train_iter.argtypes = [_vp, _vp, C.c_uint32, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, C.POINTER(C.c_uint32)]
        L.orc_object_train_iter_rng.restype = C.c_float
        L.orc_object_train_iter_rng.argtypes = [_vp, _vp, C.c_uint32, _vp, C.c_int, C.c_int, _vp, C.c_uint64]
        L.orc_object_last.restype = C.c_size_t
        L.orc_object_last.argtypes = [_vp, C.c_int, _vp, C.c_size_t]
        L.orc_object_render.argtypes = [_vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _vp, _vp, C.c_uint32, _vp, C.c_int, _vp, _vp, _vp]
        L.orc_pcg32_floats.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, _vp]
        L.orc_seed_seq_1.argtypes = [C.c_uint32, _vp]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def mat16(m) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(m, dtype=np.float32).reshape(4, 4).T).reshape(16)


def default_config(**over) -> OrcConfig:
    c = OrcConfig()
    lib().orc_default_config(C.byref(c))
    for k, v in over.items():
        setattr(c, k, v)
    return c


def grid_layout(cfg: OrcConfig):
    L = cfg.n_levels
    off = np.zeros(L + 1, np.uint32)
    sc = np.zeros(L, np.float32)
    res = np.zeros(L, np.uint32)
    n = lib().orc_grid_layout(C.byref(cfg), _p(off), _p(sc), _p(res))
    return off, sc, res, int(n)


def n_mlp_params(cfg) -> int:
    return int(lib().orc_n_mlp_params(C.byref(cfg)))


def n_params(cfg) -> int:
    return int(lib().orc_n_params(C.byref(cfg)))


def seed_seq_1(seed: int):
    out = np.zeros(2, np.uint32)
    lib().orc_seed_seq_1(seed, _p(out))
    return out


def pcg32_floats(initstate: int, advance: int, n: int) -> np.ndarray:
    out = np.zeros(n, np.float32)
    lib().orc_pcg32_floats(initstate, advance, n, _p(out))
    return out


def init_params(cfg, seed: int = 1337) -> np.ndarray:
    out = np.zeros(n_params(cfg), np.float32)
    lib().orc_init_params(C.byref(cfg), C.c_uint32(seed), _p(out))
    return out


def f2h(a) -> np.ndarray:
    a = _f32(a)
    out = np.zeros(a.shape, np.uint16)
    lib().orc_f2h_array(_p(a), _p(out), C.c_size_t(a.size))
    return out


def h2f(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint16)
    out = np.zeros(a.shape, np.float32)
    lib().orc_h2f_array(_p(a), _p(out), C.c_size_t(a.size))
    return out


def ray_intersect(bmin, bmax, o, d):
    t0, t1 = C.c_float(0), C.c_float(0)
    hit = lib().orc_ray_intersect(_p(_f32(bmin)), _p(_f32(bmax)), _p(_f32(o)), _p(_f32(d)), C.byref(t0), C.byref(t1))
    return bool(hit), t0.value, t1.value


class Frames:
    """Keeps numpy frame buffers alive and exposes them as an orc_frame array."""

    def __init__(self, rgb_list, inst_list, depth_list, poses):
        n = len(rgb_list)
        self.rgb = [np.ascontiguousarray(a, dtype=np.uint8) for a in rgb_list]
        self.inst = [np.ascontiguousarray(a, dtype=np.uint8) for a in inst_list]
        self.depth = [None if (depth_list is None or depth_list[i] is None) else _f32(depth_list[i]) for i in range(n)]
        self.arr = (OrcFrame * n)()
        for i in range(n):
            self.arr[i].rgb = self.rgb[i].ctypes.data
            self.arr[i].instance = self.inst[i].ctypes.data
            self.arr[i].depth = None if self.depth[i] is None else self.depth[i].ctypes.data
            self.arr[i].pose = (C.c_float * 16)(*mat16(poses[i]))


def make_boxes(rows):
    rows = list(rows)
    arr = (OrcBbox2d * max(1, len(rows)))()
    for i, r in enumerate(rows):
        arr[i] = OrcBbox2d(*[int(v) for v in r])
    return arr


def generate_rays(R, boxes_rows, frames: Frames, H, W, K, Tow, bmin, bmax, instance_id, use_depth, sample_xy, rand_colors):
    rows = list(boxes_rows)
    rays = np.zeros((R, 9), np.float32)
    rinst = np.zeros(R, np.uint8)
    tgt = np.zeros((R, 3), np.float32)
    tgtd = np.zeros(R, np.float32)
    n_in = lib().orc_generate_rays(C.c_uint32(R), make_boxes(rows), C.c_uint32(len(rows)), frames.arr, C.c_int(H), C.c_int(W),
                                   _p(_f32(K)), _p(mat16(Tow)), _p(_f32(bmin)), _p(_f32(bmax)), C.c_uint8(instance_id),
                                   C.c_int(int(use_depth)), _p(_f32(sample_xy)), _p(_f32(rand_colors)),
                                   _p(rays), _p(rinst), _p(tgt), _p(tgtd))
    return int(n_in), rays, rinst, tgt, tgtd


def sample_points(rays: np.ndarray, S: int, bmin, bmax, rand_dt):
    R = rays.shape[0]
    pts = np.zeros((R * S, 3), np.float32)
    t = np.zeros(R * S, np.float32)
    lib().orc_sample_points(C.c_uint32(R), C.c_uint32(S), _p(_f32(rays)), _p(_f32(bmin)), _p(_f32(bmax)), _p(_f32(rand_dt)), _p(pts), _p(t))
    return pts, t


def render_rays(box, Twc, K, Tow, bmin, bmax):
    """GenerateRenderRays: (rays [h*w][9], in_box [h*w]) for box = (FrameId, x, y, h, w)."""
    _, x, y, h, w = [int(v) for v in box]
    rays = np.zeros((h * w, 9), np.float32)
    inb = np.zeros(h * w, np.int32)
    lib().orc_render_rays(C.c_uint32(x), C.c_uint32(y), C.c_uint32(h), C.c_uint32(w), _p(mat16(Twc)), _p(_f32(K)), _p(mat16(Tow)),
                          _p(_f32(bmin)), _p(_f32(bmax)), _p(rays), _p(inb))
    return rays, inb


def volume_render_test(S2, out4, t, in_box, d_norm, bg=1.0):
    """VolumeRender_Render on raw fp32 network outputs [n_rays*S2][4]: (rgb [n][3], depth [n], mask [n])."""
    inb = np.ascontiguousarray(in_box, dtype=np.int32)
    n = inb.size
    rgb = np.zeros((n, 3), np.float32)
    dep = np.zeros(n, np.float32)
    mask = np.zeros(n, np.float32)
    lib().orc_volume_render_test(C.c_uint32(n), C.c_uint32(S2), _p(_f32(out4)), _p(_f32(t)), _p(inb), _p(_f32(d_norm)), C.c_float(bg),
                                 _p(rgb), _p(dep), _p(mask))
    return rgb, dep, mask


def encode(cfg, grid_fp16_bits, points) -> np.ndarray:
    pts = _f32(points).reshape(-1, 3)
    grid = np.ascontiguousarray(grid_fp16_bits, dtype=np.uint16)
    out = np.zeros((pts.shape[0], cfg.n_levels * 2), np.uint16)
    lib().orc_encode(C.byref(cfg), _p(grid), _p(pts), C.c_uint32(pts.shape[0]), _p(out))
    return out


def encode_corners(cfg, points):
    pts = _f32(points).reshape(-1, 3)
    idx = np.zeros((pts.shape[0], cfg.n_levels, 8), np.uint32)
    w = np.zeros((pts.shape[0], cfg.n_levels, 8), np.float32)
    lib().orc_encode_corners(C.byref(cfg), _p(pts), C.c_uint32(pts.shape[0]), _p(idx), _p(w))
    return idx, w


def mlp_forward(cfg, mlp_fp16_bits, enc_bits):
    enc = np.ascontiguousarray(enc_bits, dtype=np.uint16)
    N = enc.shape[0]
    hidden = np.zeros((cfg.n_hidden_layers, N, cfg.n_neurons), np.uint16)
    out = np.zeros((N, 16), np.uint16)
    lib().orc_mlp_forward(C.byref(cfg), _p(np.ascontiguousarray(mlp_fp16_bits, dtype=np.uint16)), _p(enc), C.c_uint32(N), _p(hidden), _p(out))
    return hidden, out


def volume_render(R, S, out_bits, t, bg):
    rgb = np.zeros((R, 3), np.float32)
    dep = np.zeros(R, np.float32)
    mask = np.zeros(R, np.float32)
    lib().orc_volume_render(C.c_uint32(R), C.c_uint32(S), _p(np.ascontiguousarray(out_bits, dtype=np.uint16)), _p(_f32(t)), _p(_f32(bg)),
                            _p(rgb), _p(dep), _p(mask))
    return rgb, dep, mask


def loss_backward(R, S, loss_scale, out_bits, t, rays_instance, target, target_depth, rgb_rays, depth_rays, mask_rays):
    dout = np.zeros((R * S, 16), np.uint16)
    loss = np.zeros(R, np.float32)
    lib().orc_loss_backward(C.c_uint32(R), C.c_uint32(S), C.c_float(loss_scale), _p(np.ascontiguousarray(out_bits, dtype=np.uint16)),
                            _p(_f32(t)), _p(np.ascontiguousarray(rays_instance, dtype=np.uint8)), _p(_f32(target)), _p(_f32(target_depth)),
                            _p(_f32(rgb_rays)), _p(_f32(depth_rays)), _p(_f32(mask_rays)), _p(dout), _p(loss))
    return dout, loss


def mlp_backward(cfg, mlp_fp16_bits, enc_bits, hidden_bits, dout_bits, round_fp16=True):
    enc = np.ascontiguousarray(enc_bits, dtype=np.uint16)
    N = enc.shape[0]
    d_enc = np.zeros((N, cfg.n_levels * 2), np.uint16)
    dW = np.zeros(n_mlp_params(cfg), np.float32)
    lib().orc_mlp_backward(C.byref(cfg), _p(np.ascontiguousarray(mlp_fp16_bits, dtype=np.uint16)), _p(enc),
                           _p(np.ascontiguousarray(hidden_bits, dtype=np.uint16)), _p(np.ascontiguousarray(dout_bits, dtype=np.uint16)),
                           C.c_uint32(N), _p(d_enc), _p(dW), C.c_int(int(round_fp16)))
    return d_enc, dW


def encode_backward(cfg, points, d_enc_bits, mode=0):
    pts = _f32(points).reshape(-1, 3)
    n_grid = n_params(cfg) - n_mlp_params(cfg)
    grad = np.zeros(n_grid, np.float32)
    lib().orc_encode_backward(C.byref(cfg), _p(pts), _p(np.ascontiguousarray(d_enc_bits, dtype=np.uint16)), C.c_uint32(pts.shape[0]), _p(grad), C.c_int(mode))
    return grad


def optimizer_step(cfg, step, grads, pf, ph, m, v, psteps, ema):
    """In-place on the numpy arrays (pf,m,v float32; ph,ema uint16; psteps uint32)."""
    lib().orc_optimizer_step(C.byref(cfg), C.c_uint32(step), _p(_f32(grads)), _p(pf), _p(ph), _p(m), _p(v), _p(psteps), _p(ema))


class OracleObject:
    STATE = {"master": 0, "params": 1, "ema": 2, "grad": 3, "adam_m": 4, "adam_v": 5, "param_steps": 6}
    LAST = {"rays": 0, "points": 1, "t": 2, "enc": 3, "out": 4, "rgb_rays": 5, "depth_rays": 6, "mask_rays": 7, "dout": 8,
            "d_enc": 9, "target": 10, "target_depth": 11, "ray_instance": 12, "loss": 13}

    def __init__(self, cfg: OrcConfig, R: int, S: int, Tow, bmin, bmax, instance_id: int, use_depth: bool, seed: int = 1337, n_threads: int = 1):
        self.cfg, self.R, self.S = cfg, R, S
        self._tow, self._bmin, self._bmax = mat16(Tow), _f32(bmin), _f32(bmax)
        self._h = lib().orc_object_create(C.byref(cfg), seed, R, S, _p(self._tow), _p(self._bmin), _p(self._bmax), instance_id, int(use_depth), n_threads)
        self.n_params = int(lib().orc_object_n_params(self._h))

    def train_iter(self, boxes_rows, frames: Frames, H, W, K, sample_xy, rand_colors, rand_dt):
        rows = list(boxes_rows)
        n_in = C.c_uint32(0)
        loss = lib().orc_object_train_iter(self._h, make_boxes(rows), len(rows), frames.arr, H, W, _p(_f32(K)),
                                           _p(_f32(sample_xy)), _p(_f32(rand_colors)), _p(_f32(rand_dt)), C.byref(n_in))
        return float(loss), int(n_in.value)

    def train_iter_rng(self, boxes_rows, frames: Frames, H, W, K, iter_seed: int) -> float:
        rows = list(boxes_rows)
        return float(lib().orc_object_train_iter_rng(self._h, make_boxes(rows), len(rows), frames.arr, H, W, _p(_f32(K)), iter_seed))

    def state(self, which: str) -> np.ndarray:
        out = np.zeros(self.n_params, np.float32)
        lib().orc_object_get(self._h, self.STATE[which], _p(out))
        return out

    def set_params(self, p):
        p = _f32(p)
        assert p.size == self.n_params
        lib().orc_object_set_params(self._h, _p(p))

    def last(self, which: str) -> np.ndarray:
        n = lib().orc_object_last(self._h, self.LAST[which], None, 0)
        out = np.zeros(n, np.float32)
        lib().orc_object_last(self._h, self.LAST[which], _p(out), n)
        return out

    def render(self, box, Twc, K, S2, rand_dt, use_ema=True):
        fid, x, y, h, w = [int(v) for v in box]
        rgb = np.zeros((h, w, 3), np.float32)
        dep = np.zeros((h, w), np.float32)
        mask = np.zeros((h, w), np.float32)
        lib().orc_object_render(self._h, x, y, h, w, _p(mat16(Twc)), _p(_f32(K)), S2, _p(_f32(rand_dt)), int(use_ema), _p(rgb), _p(dep), _p(mask))
        return rgb, dep, mask

    def close(self):
        if self._h:
            lib().orc_object_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
