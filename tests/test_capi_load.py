"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/mon_c.h declares; host-only entry points (config parsing, table geometry) behave like the
reference's ReadNetworkConfig/ResetNetwork; compute entry points fail loudly without a GPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from ro_map_b200 import build, _capi
    build.build()
    return _capi.load()


def test_every_declared_symbol_is_exported(lib):
    from ro_map_b200 import _capi
    header = (ROOT / "include" / "mon_c.h").read_text()
    declared = set(re.findall(r"\b(mon_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in mon_c.h"
    assert declared == set(_capi.SIGNATURES), declared ^ set(_capi.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layouts_match_reference_pods():
    from ro_map_b200 import _capi
    assert C.sizeof(_capi.Bbox2d) == 20          # nerf::FrameIdAndBbox, 5 x u32 (common.h:18-23)
    assert [f[0] for f in _capi.Bbox2d._fields_] == ["FrameId", "x", "y", "h", "w"]
    assert C.sizeof(_capi.Config) == 23 * 4


def test_default_config_is_base_json(lib, tmp_path):
    from ro_map_b200 import core
    cfg = core.default_config()
    assert (cfg.n_levels, cfg.n_features_per_level, cfg.log2_hashmap_size, cfg.base_resolution) == (16, 2, 16, 16)
    assert cfg.per_level_scale == 2.0 and cfg.n_neurons == 64 and cfg.n_hidden_layers == 1
    assert cfg.rays_per_batch == 4096 and cfg.samples_per_ray == 32 and cfg.render_samples_per_ray == 64
    assert core.param_counts(cfg) == (3072, 1908736)
    # the reference's own base.json text (with a comment, which its parser ignores too)
    js = tmp_path / "base.json"
    js.write_text("""{
      // comment
      "loss": {"otype": "Huber"},
      "optimizer": {"otype": "Ema", "decay": 0.95, "nested": {"otype": "ExponentialDecay", "decay_start": 20000,
         "decay_interval": 10000, "decay_base": 0.33, "nested": {"otype": "Adam", "learning_rate": 1e-2, "beta1": 0.9,
         "beta2": 0.99, "epsilon": 1e-15, "l2_reg": 1e-6}}},
      "encoding": {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 16, "base_resolution": 16},
      "network": {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 1}
    }""")
    parsed = core.config_from_json(js)
    for name, _ in cfg._fields_:
        assert getattr(parsed, name) == getattr(cfg, name), name


def test_config_errors(lib, tmp_path):
    from ro_map_b200 import core
    with pytest.raises(core.MonError, match="MON_ERR_IO"):
        core.config_from_json(tmp_path / "missing.json")
    bad = tmp_path / "bad.json"
    bad.write_text('{"encoding": {"otype": "Frequency"}, "optimizer": {"otype": "Adam"}}')
    with pytest.raises(core.MonError, match="not supported"):
        core.config_from_json(bad)
    bad.write_text('{"encoding": {"n_levels": 16, "log2_hashmap_size": 16}, "network": {"n_neurons": 64, "n_hidden_layers": 1}, "optimizer": {"otype": "Adam"')
    with pytest.raises(core.MonError, match="JSON error"):
        core.config_from_json(bad)
    # tables larger than 2^16 entries per level do not fit the shared-memory resident slices of the encode / scatter kernels:
    # rejected with a message, not an overrun (tcnn's default when the key is missing is 19)
    bad.write_text('{"encoding": {"n_levels": 16, "base_resolution": 16}, "network": {"n_neurons": 64, "n_hidden_layers": 1}, "optimizer": {"otype": "Adam"}}')
    with pytest.raises(core.MonError, match="log2_hashmap_size"):
        core.config_from_json(bad)
    cfg = core.default_config()
    cfg.log2_hashmap_size = 17
    with pytest.raises(core.MonError, match="log2_hashmap_size"):
        core.param_counts(cfg)


def test_grid_layout_matches_oracle(lib, oracle):
    from ro_map_b200 import core
    for log2, base, levels in [(16, 16, 16), (14, 16, 16), (15, 8, 16)]:
        cfg = core.default_config(log2_hashmap_size=log2, base_resolution=base, n_levels=levels)
        off, sc, res = core.grid_layout(cfg)
        ocfg = oracle.default_config(log2_hashmap_size=log2, base_resolution=base, n_levels=levels)
        o_off, o_sc, o_res, _ = oracle.grid_layout(ocfg)
        assert np.array_equal(off, o_off) and np.array_equal(sc, o_sc) and np.array_equal(res, o_res)


def test_encode_work_split_tiles_the_job_space(lib):
    """k_encode_forward cuts the flattened [(level, feature) job][point] space into one piece per CTA by a closed-form
    cost model (dense vs hashed levels, table staging per job).  Whatever the sizes, the pieces must cover every
    (job, point) exactly once — a gap would silently leave encodings stale.  Host mirror of the kernel's two functions."""
    import ctypes as C
    from ro_map_b200 import core
    rng = np.random.default_rng(5)
    cases = [(131072, 148, 0, 16), (32768, 148, 0, 16), (8192, 16, 0, 16), (1, 1, 0, 16), (7, 148, 0, 16), (1048576, 148, 0, 16),
             (131072, 148, 0, 8), (131072, 148, 8, 16), (131072, 74, 4, 12), (640000 * 64 // 39, 148, 0, 16)]
    cases += [(int(rng.integers(1, 300000)), int(rng.integers(1, 160)), 0, 16) for _ in range(40)]
    for log2, base in ((16, 16), (12, 16), (14, 8)):
        cfg = core.default_config(log2_hashmap_size=log2, base_resolution=base)
        for n, ctas, lb, le in cases:
            out = np.zeros((ctas, 4), np.uint32)
            assert lib.mon_debug_encode_pieces(C.byref(cfg), n, ctas, lb, le, out.ctypes.data_as(C.POINTER(C.c_uint32))) == 0
            cover = {j: [] for j in range(2 * lb, 2 * le)}
            for jb, pb, je, pe in out.tolist():
                job = jb
                while job < 2 * le and (job < je or (job == je and pe > 0)):      # the kernel's loop
                    p0 = pb if job == jb else 0
                    p1 = pe if job == je else n
                    if p1 > p0:
                        cover[job].append((p0, p1))
                    job += 1
            for j, segs in cover.items():
                segs.sort()
                pos = 0
                for a, b in segs:
                    assert a == pos, (log2, n, ctas, j, segs)
                    pos = b
                assert pos == n, (log2, n, ctas, j, segs)


def test_scatter_work_split_tiles_the_job_space(lib):
    """k_scatter_resident cuts the flattened [(level, index class) job][live sample] space into one piece per CTA by cost
    (coarse dense levels are dearer: same-address shared-memory atomics).  Every (job, sample) must be covered exactly once —
    a gap would silently drop gradient.  Host mirror of the kernel's two functions."""
    import ctypes as C
    from ro_map_b200 import core
    rng = np.random.default_rng(6)
    cases = [(131072, 140), (131072, 148), (24576, 140), (1, 1), (7, 140), (3, 64), (1048576, 140)]
    cases += [(int(rng.integers(1, 300000)), int(rng.integers(1, 160))) for _ in range(40)]
    for log2, base in ((16, 16), (12, 16), (14, 8)):
        cfg = core.default_config(log2_hashmap_size=log2, base_resolution=base)
        for n, ctas in cases:
            out = np.zeros((ctas, 4), np.uint32)
            assert lib.mon_debug_scatter_pieces(C.byref(cfg), n, ctas, out.ctypes.data_as(C.POINTER(C.c_uint32))) == 0
            cover = {j: [] for j in range(64)}
            for jb, pb, je, pe in out.tolist():
                job = jb
                while job < 64 and (job < je or (job == je and pe > 0)):      # the kernel's loop
                    p0 = pb if job == jb else 0
                    p1 = pe if job == je else n
                    if p1 > p0:
                        cover[job].append((p0, p1))
                    job += 1
            for j, segs in cover.items():
                segs.sort()
                pos = 0
                for a, b in segs:
                    assert a == pos, (log2, n, ctas, j, segs)
                    pos = b
                assert pos == n, (log2, n, ctas, j, segs)


def test_no_cpu_fallback(lib):
    """Without a CUDA device the compute entry points must fail, not silently compute on the host."""
    from ro_map_b200 import core
    if core.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(core.MonError, match="MON_ERR_NO_DEVICE"):
        core.Dataset(0, 100.0, 100.0, 50.0, 50.0, 100, 100, 4, False)
    cfg = core.default_config()
    with pytest.raises(core.MonError, match="MON_ERR_NO_DEVICE"):
        core.stage_encode(cfg, np.zeros(1908736, np.uint16), np.zeros((4, 3), np.float32))


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under ro_map_b200/ or include/ may mention it."""
    for p in list((ROOT / "ro_map_b200").rglob("*")) + list((ROOT / "include").rglob("*")):
        if p.is_file() and p.suffix in {".py", ".cu", ".cuh", ".h", ".cpp", ".hpp"}:
            text = p.read_text(errors="ignore")
            assert "mon_oracle" not in text and "orc_" not in text and "libmon_ref" not in text, p


def test_header_is_plain_c(tmp_path):
    """include/mon_c.h is the drop-in boundary: it must compile as C99 (no C++, no torch, no CUDA types in the signatures) so that
    cgo / JNI / ctypes-style hosts can bind it."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "mon_c.h"\nint main(void) { mon_config c; mon_bbox2d b; (void)c; (void)b; return (int)sizeof(mon_bbox2d) - 20; }\n')
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{ROOT / 'include'}", str(src), "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0          # FrameIdAndBbox layout: 5 x u32
