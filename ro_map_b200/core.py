"""Thin Python handles over the C ABI (include/mon_c.h), used by tests, bench.py and tools.

Names follow the reference's domain: a `Dataset` is nerf::NeRF_Dataset (keyframes resident on one
GPU, MON/Core/include/nerf_data.h:19-72), a `NerfObject` is nerf::NeRF + nerf::NeRF_Model (one
object's hash grid, MLP, optimizer state and batch workspace, MON/Core/include/nerf_model.h:81-185).
All arrays are numpy on the host; the device side lives entirely inside libmon_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import Bbox2d, Config, MonError, check  # noqa: F401


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None and a.size != int(np.prod(shape)):
        raise ValueError(f"expected {int(np.prod(shape))} floats, got {a.size}")
    return a


def _mat16(m) -> np.ndarray:
    """4x4 matrix (row-major numpy, math convention) -> column-major float[16], Eigen::Matrix4f::data()."""
    m = np.asarray(m, dtype=np.float32).reshape(4, 4)
    return np.ascontiguousarray(m.T).reshape(16)


def device_count() -> int:
    n = C.c_int(0)
    rc = _capi.load().mon_device_count(C.byref(n))
    if rc == -5:
        return 0
    check(rc)
    return n.value


def default_config(**over) -> Config:
    cfg = Config()
    check(_capi.load().mon_config_default(C.byref(cfg)))
    for k, v in over.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def config_from_json(path: str, **over) -> Config:
    cfg = Config()
    check(_capi.load().mon_config_from_json(str(path).encode(), C.byref(cfg)))
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


def param_counts(cfg: Config) -> tuple[int, int]:
    a, b = C.c_uint32(0), C.c_uint32(0)
    check(_capi.load().mon_config_param_counts(C.byref(cfg), C.byref(a), C.byref(b)))
    return a.value, b.value


def grid_layout(cfg: Config):
    L = cfg.n_levels
    off = (C.c_uint32 * (L + 1))()
    sc = (C.c_float * L)()
    res = (C.c_uint32 * L)()
    check(_capi.load().mon_config_grid_layout(C.byref(cfg), off, sc, res))
    return np.array(off[:], dtype=np.uint32), np.array(sc[:], dtype=np.float32), np.array(res[:], dtype=np.uint32)


def make_boxes(rows) -> "C.Array[Bbox2d]":
    """rows: iterable of (FrameId, x, y, h, w) — the reference's file order (nerf.cu:67-113)."""
    rows = list(rows)
    arr = (Bbox2d * max(1, len(rows)))()
    for i, r in enumerate(rows):
        arr[i] = Bbox2d(*[int(v) for v in r])
    return arr


def _read_mesh(lib, h) -> dict:
    try:
        nv, ns, ni = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        check(lib.mon_mesh_counts(h, C.byref(nv), C.byref(ns), C.byref(ni)))
        verts, normals = np.zeros((nv.value, 3), np.float32), np.zeros((nv.value, 3), np.float32)
        colors, idx = np.zeros((nv.value, 3), np.uint8), np.zeros(ni.value, np.uint32)
        check(lib.mon_mesh_read(h, _ptr(verts), _ptr(normals), _ptr(colors), _ptr(idx)))
        return {"verts": verts, "normals": normals, "colors": colors, "indices": idx, "n_surface": ns.value}
    finally:
        lib.mon_mesh_destroy(h)


def mesh_from_lattice(sigma_zyx: np.ndarray, bmin, bmax, thresh: float = 2.0, gpu: int = 0) -> dict:
    """Marching cubes of the core (csrc/kernels_mesh.cu) on a caller's cubic lattice [z][y][x]."""
    lib = _capi.load()
    sig = np.ascontiguousarray(sigma_zyx, dtype=np.float32)
    res = sig.shape[0]
    if sig.shape != (res, res, res):
        raise ValueError("cubic lattice expected")
    lo, hi = _f32(bmin, (3,)), _f32(bmax, (3,))
    h = C.c_void_p()
    check(lib.mon_mesh_from_lattice(gpu, _ptr(sig), res, lo.ctypes.data_as(C.POINTER(C.c_float)), hi.ctypes.data_as(C.POINTER(C.c_float)), float(thresh), C.byref(h)))
    return _read_mesh(lib, h)


class Dataset:
    def __init__(self, gpu: int, fx: float, fy: float, cx: float, cy: float, H: int, W: int, max_frames: int, use_depth: bool):
        self._lib = _capi.load()
        self._h = C.c_void_p()
        self.H, self.W, self.K = H, W, (fx, fy, cx, cy)
        self.use_depth = bool(use_depth)
        self.gpu = gpu
        check(self._lib.mon_dataset_create(gpu, fx, fy, cx, cy, H, W, max_frames, int(use_depth), C.byref(self._h)))
        self.depth_u16 = False

    def set_depth_u16(self, depth_factor: float):
        """From here on the depth planes are the raw 16-bit samples of the depth image (uint16 blocks); the batch kernel converts the
        pixels it picks, (float)u16 * depth_factor.  Before the first keyframe only."""
        check(self._lib.mon_dataset_set_depth_u16(self._h, float(depth_factor)))
        self.depth_u16 = True

    def _depth_dtype(self):
        return np.uint16 if self.depth_u16 else np.float32

    def add_frame(self, frame_id: int, rgb_u8: np.ndarray, instance_u8: np.ndarray, depth, pose_c2w, is_bgr: bool = False):
        """depth: HxW float32 metres, or HxW uint16 counts after set_depth_u16()."""
        rgb = np.ascontiguousarray(rgb_u8, dtype=np.uint8)
        inst = np.ascontiguousarray(instance_u8, dtype=np.uint8)
        if rgb.size != self.H * self.W * 3 or inst.size != self.H * self.W:
            raise ValueError("frame has the wrong size")
        d = None
        if depth is not None:
            if self.depth_u16 and np.asarray(depth).dtype != np.uint16:
                raise ValueError("the dataset takes raw uint16 depth")
            d = np.ascontiguousarray(depth, dtype=self._depth_dtype())
            if d.size != self.H * self.W:
                raise ValueError("depth plane has the wrong size")
        pose = _mat16(pose_c2w)
        fn = self._lib.mon_dataset_add_frame_d16 if self.depth_u16 else self._lib.mon_dataset_add_frame
        check(fn(self._h, frame_id, _ptr(rgb), int(is_bgr), _ptr(inst), None if d is None else _ptr(d), pose.ctypes.data_as(C.POINTER(C.c_float))))

    def add_frame_device(self, frame_id: int, d_rgb: int, d_instance: int, d_depth, pose_c2w, is_bgr: bool = False):
        """Keyframe planes already in this GPU's memory (raw device addresses, e.g. torch.Tensor.data_ptr())."""
        pose = _mat16(pose_c2w)
        fn = self._lib.mon_dataset_add_frame_device_d16 if self.depth_u16 else self._lib.mon_dataset_add_frame_device
        check(fn(self._h, frame_id, C.c_void_p(d_rgb), int(is_bgr), C.c_void_p(d_instance),
                 None if d_depth is None else C.c_void_p(d_depth), pose.ctypes.data_as(C.POINTER(C.c_float))))

    def add_frames(self, first_id: int, rgb_u8, instance_u8, depth_f32, poses_c2w, is_bgr: bool = False):
        """n consecutive keyframes in one call.  rgb_u8 [n, H, W, 3], instance_u8 [n, H, W], depth_f32 [n, H, W] (uint16 counts after
        set_depth_u16()) or None: numpy blocks
        (page-locked ones are DMA-ed straight out of, three copies per slab of 32 frames) or raw device addresses (ints, blocks in
        this GPU's memory; the frame count then comes from poses_c2w)."""
        n = len(poses_c2w)
        flat = np.concatenate([_mat16(p) for p in poses_c2w]).astype(np.float32)
        on_device = isinstance(rgb_u8, int)
        if on_device:
            a_rgb, a_inst, a_dep = C.c_void_p(rgb_u8), C.c_void_p(instance_u8), None if depth_f32 is None else C.c_void_p(depth_f32)
        else:
            for a, per in ((rgb_u8, self.H * self.W * 3), (instance_u8, self.H * self.W)):
                if a.dtype != np.uint8 or not a.flags.c_contiguous or a.size != n * per:
                    raise ValueError("frame block has the wrong size, dtype or layout")
            if depth_f32 is not None and (depth_f32.dtype != self._depth_dtype() or not depth_f32.flags.c_contiguous or depth_f32.size != n * self.H * self.W):
                raise ValueError("depth block has the wrong size, dtype or layout")
            a_rgb, a_inst, a_dep = _ptr(rgb_u8), _ptr(instance_u8), None if depth_f32 is None else _ptr(depth_f32)
        fn = self._lib.mon_dataset_add_frames_d16 if self.depth_u16 else self._lib.mon_dataset_add_frames
        check(fn(self._h, first_id, n, a_rgb, int(is_bgr), a_inst, a_dep, flat.ctypes.data_as(C.POINTER(C.c_float)), int(on_device)))

    def sync(self):
        check(self._lib.mon_dataset_sync(self._h))

    def update_poses(self, first: int, poses_c2w):
        flat = np.concatenate([_mat16(p) for p in poses_c2w]).astype(np.float32)
        check(self._lib.mon_dataset_update_poses(self._h, first, len(poses_c2w), flat.ctypes.data_as(C.POINTER(C.c_float))))

    @property
    def frame_count(self) -> int:
        n = C.c_uint32(0)
        check(self._lib.mon_dataset_frame_count(self._h, C.byref(n)))
        return n.value

    def clone_from_peer(self, src: "Dataset"):
        check(self._lib.mon_dataset_clone_from_peer(self._h, src._h))

    def copy_frame_from_peer(self, src: "Dataset", frame_id: int):
        """One keyframe of `src`, device to device, asynchronously on this dataset's upload stream (online replication)."""
        check(self._lib.mon_dataset_copy_frame_from_peer(self._h, src._h, frame_id))

    def close(self):
        if self._h:
            self._lib.mon_dataset_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NerfObject:
    STATE = {"master": 0, "params": 1, "ema": 2, "grad": 3, "adam_m": 4, "adam_v": 5, "param_steps": 6}
    LAST = {"rays": 0, "points": 1, "enc": 3, "out": 4, "rgb_rays": 5, "depth_rays": 6, "mask_rays": 7, "dout": 8, "d_enc": 9,
            "target": 10, "target_depth": 11, "ray_instance": 12, "loss": 13}

    def __init__(self, ds: Dataset, cfg: Config, obj_Tow, bmin, bmax, instance_id: int, seed: int = 1337):
        self._lib = _capi.load()
        self._h = C.c_void_p()
        self.ds = ds  # keeps the dataset alive
        self.cfg = cfg
        self.R, self.S = cfg.rays_per_batch, cfg.samples_per_ray
        self.n_mlp, self.n_grid = param_counts(cfg)
        self.n_params = self.n_mlp + self.n_grid
        tow = _mat16(obj_Tow)
        bmin, bmax = _f32(bmin, (3,)), _f32(bmax, (3,))
        fp = C.POINTER(C.c_float)
        check(self._lib.mon_object_create(ds._h, C.byref(cfg), seed, instance_id, tow.ctypes.data_as(fp),
                                           bmin.ctypes.data_as(fp), bmax.ctypes.data_as(fp), C.byref(self._h)))

    # ---- NeRF_Model::UpdateFrameIdAndBbox / ...Online
    def set_bboxes(self, rows):
        rows = list(rows)
        check(self._lib.mon_object_set_bboxes(self._h, make_boxes(rows), len(rows)))

    def add_bboxes(self, rows):
        rows = list(rows)
        check(self._lib.mon_object_add_bboxes(self._h, make_boxes(rows), len(rows)))

    # ---- NeRF_Model::Train_Step
    def train(self, iters: int) -> float:
        loss = C.c_float(0)
        check(self._lib.mon_object_train(self._h, iters, C.byref(loss)))
        return loss.value

    def train_async(self, iters: int):
        check(self._lib.mon_object_train_async(self._h, iters))

    def prepare_train(self, iters: int):
        """Capture the iteration graphs of a call of `iters` iterations now (AllocateBatchWorkspace's role)."""
        check(self._lib.mon_object_prepare_train(self._h, iters))

    def sync(self):
        check(self._lib.mon_object_sync(self._h))

    STAGES = ("batch", "points", "encode", "mlp_fused", "scatter", "optimizer")

    def train_profiled(self, iters: int) -> dict:
        """Mean device milliseconds per stage (CUDA events on the object's stream), see mon_c.h."""
        ms = (C.c_float * len(self.STAGES))()
        check(self._lib.mon_object_train_profiled(self._h, iters, ms, len(self.STAGES)))
        return dict(zip(self.STAGES, [float(x) for x in ms]))

    @property
    def last_train_ms(self) -> float:
        ms = C.c_float(0)
        check(self._lib.mon_object_last_train_ms(self._h, C.byref(ms)))
        return ms.value

    @property
    def step(self) -> int:
        s = C.c_uint32(0)
        check(self._lib.mon_object_step_count(self._h, C.byref(s)))
        return s.value

    @property
    def live_fraction(self) -> float:
        """Share of the last iteration's samples that carried gradient (work of the scatter + Adam kernel)."""
        a, b = C.c_uint32(0), C.c_uint32(0)
        check(self._lib.mon_object_live_samples(self._h, C.byref(a), C.byref(b)))
        return a.value / max(1, b.value)

    @property
    def launch_count(self) -> int:
        n = C.c_uint64(0)
        check(self._lib.mon_object_launch_count(self._h, C.byref(n)))
        return n.value

    # ---- NeRF_Model::Render
    def render(self, box, Twc, use_ema: bool = True, rand_dt=None, object_centric: bool = False):
        """object_centric: Twc is a camera -> OBJECT pose (one view of RenderVideo, nerf_model.cu:1832-1991)."""
        fid, x, y, h, w = [int(v) for v in box]
        n = h * w
        rgb = np.empty((h, w, 3), np.float32)
        depth = np.empty((h, w), np.float32)
        mask = np.empty((h, w), np.float32)
        twc = _mat16(Twc)
        jit = None if rand_dt is None else _f32(rand_dt, (n, self.cfg.render_samples_per_ray))
        fn = self._lib.mon_object_render_object_centric if object_centric else self._lib.mon_object_render
        check(fn(self._h, Bbox2d(fid, x, y, h, w), twc.ctypes.data_as(C.POINTER(C.c_float)), int(use_ema),
                 None if jit is None else _ptr(jit), _ptr(rgb), _ptr(depth), _ptr(mask)))
        return rgb, depth, mask

    def density_grid(self, res=(64, 64, 64)) -> np.ndarray:
        r = (C.c_uint32 * 3)(*[int(v) for v in res])
        out = np.empty((res[2], res[1], res[0]), np.float32)  # x fastest
        check(self._lib.mon_object_density_grid(self._h, r, _ptr(out)))
        return out

    def extract_mesh(self, res: int = 64, thresh: float = 2.0) -> dict:
        """GenerateMesh on the GPU: verts / normals [n, 3] (n padded to a multiple of 128), u8 colors [n, 3], indices, n_surface."""
        h = C.c_void_p()
        check(self._lib.mon_object_extract_mesh(self._h, int(res), float(thresh), C.byref(h)))
        return _read_mesh(self._lib, h)

    def query_points(self, points_unit, use_ema: bool = True) -> np.ndarray:
        """Network logits (r, g, b, sigma) at unit-cube positions [n, 3]."""
        pts = _f32(points_unit).reshape(-1, 3)
        out = np.empty((pts.shape[0], 4), np.float32)
        check(self._lib.mon_object_query_points(self._h, _ptr(pts), pts.shape[0], int(use_ema), _ptr(out)))
        return out

    # ---- parity hooks
    def train_injected(self, sample_xy, rand_colors, rand_dt) -> tuple[float, int]:
        sxy = _f32(sample_xy, (self.R, 2))
        col = _f32(rand_colors, (self.R, 3))
        dt = _f32(rand_dt, (self.R, self.S))
        loss, n_in = C.c_float(0), C.c_uint32(0)
        check(self._lib.mon_object_train_injected(self._h, _ptr(sxy), _ptr(col), _ptr(dt), C.byref(loss), C.byref(n_in)))
        return loss.value, n_in.value

    def set_occupancy(self, grid_res: int, warmup_iters: int = 256, update_interval: int = 16, alpha_threshold: float = 0.01):
        """OPT-IN occupancy grid + warp-ballot sample compaction (changes results; grid_res = 0 switches it off again)."""
        check(self._lib.mon_object_set_occupancy(self._h, grid_res, warmup_iters, update_interval, alpha_threshold))

    def occupancy_stats(self):
        a, b = C.c_float(0), C.c_float(0)
        check(self._lib.mon_object_occupancy_stats(self._h, C.byref(a), C.byref(b)))
        return {"occupied_cell_fraction": a.value, "occupied_sample_fraction": b.value}

    def state(self, which: str) -> np.ndarray:
        out = np.empty(self.n_params, np.float32)
        check(self._lib.mon_object_get_state(self._h, self.STATE[which], _ptr(out), out.size))
        return out

    def set_params(self, params_fp32):
        p = _f32(params_fp32, (self.n_params,))
        check(self._lib.mon_object_set_params(self._h, _ptr(p), p.size))

    def last(self, which: str) -> np.ndarray:
        R, N = self.R, self.R * self.S
        sizes = {"rays": R * 9, "points": N * 3, "enc": N * 32, "out": N * 4, "rgb_rays": R * 3, "depth_rays": R, "mask_rays": R, "dout": N * 4,
                 "d_enc": N * 32, "target": R * 3, "target_depth": R, "ray_instance": R, "loss": R}
        out = np.empty(sizes[which], np.float32)
        n = C.c_size_t(0)
        check(self._lib.mon_object_last(self._h, self.LAST[which], _ptr(out), out.size, C.byref(n)))
        return out[: n.value]

    def close(self):
        if self._h:
            self._lib.mon_object_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def stage_encode(cfg: Config, grid_fp16_bits: np.ndarray, points_unit: np.ndarray) -> np.ndarray:
    """Kernel-level hook: hash-grid encode of explicit unit-cube positions. Returns [N, 2L] fp16 bit patterns."""
    grid = np.ascontiguousarray(grid_fp16_bits, dtype=np.uint16)
    pts = _f32(points_unit).reshape(-1, 3)
    out = np.empty((pts.shape[0], 2 * cfg.n_levels), np.uint16)
    check(_capi.load().mon_stage_encode(C.byref(cfg), _ptr(grid), grid.size, _ptr(pts), pts.shape[0], _ptr(out)))
    return out
