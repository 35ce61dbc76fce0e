"""ctypes binding of libmon_b200.so — exactly the entry points declared in include/mon_c.h.

The library is the product; there is no Python or CPU fallback.  Importing this module never
builds anything: `ro_map_b200.build.build()` (or `__graft_entry__.build()`) produces the .so.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libmon_b200.so"

MON_OK = 0
ERR_NAMES = {-1: "MON_ERR_CUDA", -2: "MON_ERR_ARG", -3: "MON_ERR_IO", -4: "MON_ERR_STATE", -5: "MON_ERR_NO_DEVICE"}


class MonError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class Bbox2d(C.Structure):
    """== nerf::FrameIdAndBbox (MON/Core/include/common.h:18-23); field order x, y, h, w."""
    _fields_ = [("FrameId", C.c_uint32), ("x", C.c_uint32), ("y", C.c_uint32), ("h", C.c_uint32), ("w", C.c_uint32)]


class Config(C.Structure):
    _fields_ = [
        ("n_levels", C.c_uint32), ("n_features_per_level", C.c_uint32), ("log2_hashmap_size", C.c_uint32),
        ("base_resolution", C.c_uint32), ("per_level_scale", C.c_float),
        ("n_neurons", C.c_uint32), ("n_hidden_layers", C.c_uint32),
        ("learning_rate", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("epsilon", C.c_float), ("l2_reg", C.c_float),
        ("ema_decay", C.c_float),
        ("decay_start", C.c_uint32), ("decay_interval", C.c_uint32), ("decay_base", C.c_float),
        ("loss_scale", C.c_float),
        ("rays_per_batch", C.c_uint32), ("samples_per_ray", C.c_uint32), ("render_samples_per_ray", C.c_uint32),
        ("depth_lambda", C.c_float), ("mask_lambda", C.c_float), ("bg_density_reg", C.c_float),
    ]


_P = C.POINTER
_f32p = _P(C.c_float)
_u8p = _P(C.c_uint8)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/mon_c.h declares
SIGNATURES = {
    "mon_last_error": (C.c_char_p, []),
    "mon_version": (C.c_char_p, []),
    "mon_device_count": (C.c_int, [_P(C.c_int)]),
    "mon_config_default": (C.c_int, [_P(Config)]),
    "mon_config_from_json": (C.c_int, [C.c_char_p, _P(Config)]),
    "mon_config_param_counts": (C.c_int, [_P(Config), _P(C.c_uint32), _P(C.c_uint32)]),
    "mon_config_grid_layout": (C.c_int, [_P(Config), _P(C.c_uint32), _f32p, _P(C.c_uint32)]),
    "mon_dataset_create": (C.c_int, [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_uint32, C.c_int, _P(_vp)]),
    "mon_dataset_add_frame": (C.c_int, [_vp, C.c_uint32, _vp, C.c_int, _vp, _vp, _f32p]),
    "mon_dataset_add_frame_device": (C.c_int, [_vp, C.c_uint32, _vp, C.c_int, _vp, _vp, _f32p]),
    "mon_dataset_add_frames": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, C.c_int, _vp, _vp, _f32p, C.c_int]),
    "mon_dataset_set_depth_u16": (C.c_int, [_vp, C.c_float]),
    "mon_dataset_add_frame_d16": (C.c_int, [_vp, C.c_uint32, _vp, C.c_int, _vp, _vp, _f32p]),
    "mon_dataset_add_frame_device_d16": (C.c_int, [_vp, C.c_uint32, _vp, C.c_int, _vp, _vp, _f32p]),
    "mon_dataset_add_frames_d16": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, C.c_int, _vp, _vp, _f32p, C.c_int]),
    "mon_dataset_sync": (C.c_int, [_vp]),
    "mon_dataset_update_poses": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _f32p]),
    "mon_dataset_frame_count": (C.c_int, [_vp, _P(C.c_uint32)]),
    "mon_dataset_clone_from_peer": (C.c_int, [_vp, _vp]),
    "mon_dataset_copy_frame_from_peer": (C.c_int, [_vp, _vp, C.c_uint32]),
    "mon_dataset_destroy": (C.c_int, [_vp]),
    "mon_object_create": (C.c_int, [_vp, _P(Config), C.c_uint32, C.c_uint8, _f32p, _f32p, _f32p, _P(_vp)]),
    "mon_object_destroy": (C.c_int, [_vp]),
    "mon_object_set_bboxes": (C.c_int, [_vp, _P(Bbox2d), C.c_uint32]),
    "mon_object_add_bboxes": (C.c_int, [_vp, _P(Bbox2d), C.c_uint32]),
    "mon_object_train": (C.c_int, [_vp, C.c_uint32, _f32p]),
    "mon_object_train_async": (C.c_int, [_vp, C.c_uint32]),
    "mon_object_prepare_train": (C.c_int, [_vp, C.c_uint32]),
    "mon_object_live_samples": (C.c_int, [_vp, _P(C.c_uint32), _P(C.c_uint32)]),
    "mon_object_sync": (C.c_int, [_vp]),
    "mon_object_train_profiled": (C.c_int, [_vp, C.c_uint32, _f32p, C.c_uint32]),
    "mon_object_last_train_ms": (C.c_int, [_vp, _f32p]),
    "mon_object_step_count": (C.c_int, [_vp, _P(C.c_uint32)]),
    "mon_object_launch_count": (C.c_int, [_vp, _P(C.c_uint64)]),
    "mon_object_render": (C.c_int, [_vp, Bbox2d, _f32p, C.c_int, _vp, _vp, _vp, _vp]),
    "mon_debug_encode_pieces": (C.c_int, [_P(Config), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _P(C.c_uint32)]),
    "mon_object_set_occupancy": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float]),
    "mon_object_occupancy_stats": (C.c_int, [_vp, _f32p, _f32p]),
    "mon_debug_scatter_pieces": (C.c_int, [_P(Config), C.c_uint32, C.c_uint32, _P(C.c_uint32)]),
    "mon_object_render_object_centric": (C.c_int, [_vp, Bbox2d, _f32p, C.c_int, _vp, _vp, _vp, _vp]),
    "mon_object_density_grid": (C.c_int, [_vp, _P(C.c_uint32), _vp]),
    "mon_object_query_points": (C.c_int, [_vp, _vp, C.c_uint32, C.c_int, _vp]),
    "mon_object_extract_mesh": (C.c_int, [_vp, C.c_uint32, C.c_float, _P(_vp)]),
    "mon_mesh_from_lattice": (C.c_int, [C.c_int, _vp, C.c_uint32, _f32p, _f32p, C.c_float, _P(_vp)]),
    "mon_mesh_counts": (C.c_int, [_vp, _P(C.c_uint32), _P(C.c_uint32), _P(C.c_uint32)]),
    "mon_mesh_read": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "mon_mesh_destroy": (C.c_int, [_vp]),
    "mon_object_train_injected": (C.c_int, [_vp, _vp, _vp, _vp, _f32p, _P(C.c_uint32)]),
    "mon_object_get_state": (C.c_int, [_vp, C.c_int, _vp, C.c_size_t]),
    "mon_object_set_params": (C.c_int, [_vp, _vp, C.c_size_t]),
    "mon_object_last": (C.c_int, [_vp, C.c_int, _vp, C.c_size_t, _P(C.c_size_t)]),
    "mon_stage_encode": (C.c_int, [_P(Config), _vp, C.c_size_t, _vp, C.c_uint32, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Loads the CUDA core.  Raises (never falls back) when the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: run `python -m ro_map_b200.build` (nvcc, sm_100a). "
                          "ro_map_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != MON_OK:
        raise MonError(rc, load().mon_last_error().decode(errors="replace"))
