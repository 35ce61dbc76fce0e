#!/usr/bin/env python
"""Recovers the reference's marching-cubes triangle table from the OUTPUT of the reference's own marching_cubes.cu
(tests/golden/romap_mesh_golden.npz: the reference run on a B200 on a white-noise lattice in which every one of the 256 cell
configurations occurs, oracle/ref/make_golden_romap.py) and writes it as ro_map_b200/host/mc_table.h.

gen_faces (MON/Core/src/marching_cubes.cu:372-435) emits, per cell, the triangles of triangle_table[mask] in table order as one
contiguous run of the index array (one atomicAdd per cell); cells come in any order.  Every golden triangle is assigned to the
cell that holds its three vertices among its 12 edges, the runs give the order inside a configuration, and the local edge
numbering is gen_faces' own (:393-421).  Every configuration must be recovered identically from all of its cells.

    python tools/derive_mc_table.py            # rewrites ro_map_b200/host/mc_table.h, prints the checks
"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "ref"))
import make_golden_romap as mg  # noqa: E402

THRESH = np.float32(2.0)
# lattice point (dx, dy, dz) and axis of the 12 local edges, gen_faces' numbering (:405-421)
EDGE_OWNER = [((0, 0, 0), 0), ((1, 0, 0), 1), ((0, 1, 0), 0), ((0, 0, 0), 1),
              ((0, 0, 1), 0), ((1, 0, 1), 1), ((0, 1, 1), 0), ((0, 0, 1), 1),
              ((0, 0, 0), 2), ((1, 0, 0), 2), ((1, 1, 0), 2), ((0, 1, 0), 2)]
CORNERS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]   # mask bits (:391-400)


def edge_vertices(d, res):
    """position of the vertex on every sign-changing lattice edge, gen_vertices' arithmetic (:41-91) in float32"""
    lo, hi = mg.MESH_BOX
    scale = ((hi - lo) / np.float32(res - 1)).astype(np.float32)
    out = {}
    for z in range(res):
        for y in range(res):
            for x in range(res):
                f0 = d[z, y, x]
                c = (x, y, z)
                for a in range(3):
                    n = [x, y, z]
                    n[a] += 1
                    if n[a] >= res:
                        continue
                    f1 = d[n[2], n[1], n[0]]
                    if (f0 > THRESH) == (f1 > THRESH):
                        continue
                    dt = np.float32((THRESH - f0) / np.float32(f1 - f0))
                    p = [np.float32(c[k]) + (dt if k == a else np.float32(0)) for k in range(3)]
                    # fmaf(p, scale, min): one rounding; float64 holds the exact product + sum of float32 operands up to 1 rounding
                    pos = tuple(np.float32(np.float64(p[k]) * np.float64(scale[k]) + np.float64(lo[k])) for k in range(3))
                    out[(x, y, z, a)] = pos
    return out


def recover(kind, res, table):
    gold = np.load(ROOT / "tests" / "golden" / "romap_mesh_golden.npz")
    verts, idx = gold[kind + "_verts"], gold[kind + "_indices"].astype(np.int64)
    d = mg.mesh_lattice(kind, res)
    by_pos = {tuple(v): i for i, v in enumerate(verts)}
    ev = edge_vertices(d, res)
    vid = {}
    for key, pos in ev.items():
        assert pos in by_pos, (key, pos)           # the vertex sets are bit-identical
        vid[key] = by_pos[pos]
    cells_of_vertex = {}
    cell_edges, cell_mask = {}, {}
    for z in range(res - 1):
        for y in range(res - 1):
            for x in range(res - 1):
                mask = 0
                for b, (dx, dy, dz) in enumerate(CORNERS):
                    if d[z + dz, y + dy, x + dx] > THRESH:
                        mask |= 1 << b
                if mask in (0, 255):
                    continue
                le = {}
                for e, ((dx, dy, dz), a) in enumerate(EDGE_OWNER):
                    v = vid.get((x + dx, y + dy, z + dz, a))
                    if v is not None:
                        le[v] = e
                        cells_of_vertex.setdefault(v, set()).add((x, y, z))
                cell_edges[(x, y, z)], cell_mask[(x, y, z)] = le, mask
    tris = idx.reshape(-1, 3)
    cand = [set.intersection(*(cells_of_vertex[int(v)] for v in t)) for t in tris]
    assert all(cand), "a golden triangle does not lie in one cell"
    owner = [next(iter(c)) if len(c) == 1 else None for c in cand]
    # a triangle whose three edges lie in a face shared by two cells: it belongs to the run (= cell) of a neighbour in the array
    for _ in range(4):
        for i, o in enumerate(owner):
            if o is None:
                for j in (i - 1, i + 1):
                    if 0 <= j < len(owner) and owner[j] is not None and owner[j] in cand[i]:
                        owner[i] = owner[j]
                        break
    assert all(o is not None for o in owner)
    runs = {}
    for i, o in enumerate(owner):
        runs.setdefault(o, []).append(i)
    n_checked = 0
    for cell, ids in runs.items():
        assert ids == list(range(ids[0], ids[0] + len(ids))), "a cell's triangles are not one contiguous run"
        le = cell_edges[cell]
        lst = [le[int(v)] for i in ids for v in tris[i]]
        mask = cell_mask[cell]
        if mask in table:
            assert table[mask] == lst, (mask, table[mask], lst)
            n_checked += 1
        else:
            table[mask] = lst
    # cells with a surface configuration but no triangle would mean a configuration with an empty list: none exists
    assert set(runs) == set(cell_mask), "a surface cell without triangles"
    return len(runs), n_checked


def main():
    table = {}
    for kind, res in (("cases", 22), ("noise", 14), ("sphere", 24)):
        n_cells, n_checked = recover(kind, res, table)
        print(f"{kind}: {n_cells} surface cells, {n_checked} re-confirmed an already recovered configuration")
    missing = [m for m in range(1, 255) if m not in table]
    assert not missing, f"configurations not in the golden lattices: {missing}"
    rows = []
    for m in range(256):
        lst = table.get(m, [])
        assert len(lst) % 3 == 0 and len(lst) <= 15
        rows.append(lst + [-1] * (16 - len(lst)))
    out = ROOT / "ro_map_b200" / "host" / "mc_table.h"
    with open(out, "w") as f:
        f.write("// mc_table.h — GENERATED by tools/derive_mc_table.py, do not edit.\n"
                "// The 256-configuration triangle table of the reference's marching cubes (the classic published table: each row lists\n"
                "// cell-edge numbers, three per triangle, -1 terminated), recovered from the OUTPUT of the reference's own\n"
                "// marching_cubes.cu run on a B200 (tests/golden/romap_mesh_golden.npz): every configuration occurs in the white-noise\n"
                "// lattice and was recovered identically from each of its cells.  Corner / edge numbering: MON/Core/src/marching_cubes.cu:391-421.\n"
                "#pragma once\n#include <cstdint>\nnamespace mesh { namespace mc {\nstatic const int8_t TRIANGLES[256][16] = {\n")
        for m, r in enumerate(rows):
            f.write("    {" + ", ".join(f"{v:2d}" for v in r) + "}" + ("," if m < 255 else "") + "\n")
        f.write("};\n} }  // namespace mesh::mc\n")
    print(f"wrote {out}: {sum(len(table.get(m, [])) for m in range(256)) // 3} triangles over 254 surface configurations")
    # where the reference tree is present (the build container), confirm against its table without copying it
    ref = Path("/root/reference/dependencies/Multi-Object-NeRF/Core/src/marching_cubes.cu")
    if ref.exists():
        import re
        src = ref.read_text()
        body = src[src.index("triangle_table[256][16]"):]
        body = body[body.index("{") + 1: body.index("};")]
        ref_rows = [[int(v) for v in re.findall(r"-?\d+", row)] for row in re.findall(r"\{([^{}]*)\}", body)]
        same = len(ref_rows) == 256 and all(list(a) == list(b) for a, b in zip(ref_rows, rows))
        print("identical to the reference's triangle_table:", same)
        assert same


if __name__ == "__main__":
    main()
