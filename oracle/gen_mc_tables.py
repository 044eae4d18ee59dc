"""Writes voxblox_b200/csrc/vbx_mc_tables.h: the marching-cubes case table in the cube-corner /
edge numbering voxblox uses (MarchingCubes::kTriangleTable / kEdgeIndexPairs,
src/mesh/marching_cubes.cc:33-293 -- the Lorensen-Cline / Bourke table; the numbering is part of
the interface because Mesh vertex ORDER must match).  The values are read out of the reference
library built by oracle/Makefile (oracle/_ref/libvbx_ref.so, vbo_mc_tables) and packed: one
64-bit word per case, nibble k = k-th edge index of the row, 0xF = end of row.
tests/test_mesh_cpu.py pins the committed header to the reference's table.

    python oracle/gen_mc_tables.py
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "voxblox_b200", "csrc", "vbx_mc_tables.h")


def reference_tables():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libvbx_ref.so"))
    tri = (C.c_int32 * (256 * 16))()
    edges = (C.c_int32 * 24)()
    assert lib.vbo_mc_tables(tri, edges) == 0
    return [list(tri[16 * i:16 * i + 16]) for i in range(256)], [list(edges[2 * i:2 * i + 2]) for i in range(12)]


def pack_row(row):
    word = 0
    for k in range(16):
        v = row[k] if row[k] >= 0 else 0xF
        assert 0 <= v <= 0xF
        word |= v << (4 * k)
    return word


def main():
    tri, edges = reference_tables()
    for row in tri:  # every row: edge triples, then -1 to the end
        n = row.index(-1)
        assert n % 3 == 0 and n <= 15 and all(v == -1 for v in row[n:]) and all(0 <= v < 12 for v in row[:n])
    with open(OUT, "w") as f:
        f.write("// GENERATED (gen_mc_tables.py, kept with the test infrastructure) -- do not edit.\n"
                "// Marching-cubes case table in voxblox's corner / edge numbering (MarchingCubes::kTriangleTable,\n"
                "// kEdgeIndexPairs; voxblox/src/mesh/marching_cubes.cc:33-293), packed: one 64-bit word per case,\n"
                "// nibble k = k-th edge index of the row, 0xF = end of row.  Pinned to the reference's table by\n"
                "// tests/test_mesh_cpu.py.\n#pragma once\n#include <stdint.h>\n\n"
                "#define VBX_MC_TRIANGLE_WORDS \\\n")
        words = [pack_row(r) for r in tri]
        for i in range(0, 256, 4):
            f.write("  " + ", ".join("0x%016xull" % w for w in words[i:i + 4]) + (", \\\n" if i < 252 else "\n"))
        f.write("\n// the two cube corners each edge joins\n#define VBX_MC_EDGE_PAIRS \\\n  ")
        f.write(", ".join("{%d, %d}" % (a, b) for a, b in edges) + "\n")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
