#!/usr/bin/env python3
"""Regenerate the barycentric sample tables and emit them as a C header.

The reference ships its tetrahedron sample points as 6-decimal literals
(DEMO/tet_dis_coord.cpp:12-2066, levels n = 1..9 with n^3 points each) that were printed
by its own generator image3d::generate_sample_point (DEMO/image3d.cpp:256-306, "%f").
This script restates that generator -- unit cube split in n^3 cells, 6 tets per cell,
keep the tets whose centroid has x+y+z < 1, print the centroid's barycentric coordinates
w.r.t. (0,0,0),(0,1,0),(1,0,0),(0,0,1) with "%f" -- and parses the printed text back, so
the doubles are the ones a C compiler produces from the reference's literals.  The
triangle table (DEMO/tet_dis_coord.cpp:2068-3127, generator image3d.cpp:308-348) is
regenerated the same way.  Gauss quadrature sets (DEMO/gaussian_points.h:24-81) have no
generator; they are numeric constants of the algorithm and are listed here as data.

`--check REF` parses the reference's literals and verifies equality (run in the build
container where /root/reference exists; tests/test_tables.py pins a checksum for the
GPU box where it does not).

usage: gen_tables.py OUT.h [--check /root/reference]
"""
import hashlib
import re
import struct
import sys


def sub(a, b):
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


def add(a, b):
    return (a[0] + b[0], a[1] + b[1], a[2] + b[2])


def cross(x, y):
    return (x[1] * y[2] - x[2] * y[1], x[2] * y[0] - x[0] * y[2], x[0] * y[1] - x[1] * y[0])


def dot(a, b):
    return ((0.0 + a[0] * b[0]) + a[1] * b[1]) + a[2] * b[2]


def signed_volume(a, b, c, d):  # is_mesh/util.h:86-90
    return dot(sub(a, d), cross(sub(b, d), sub(c, d))) / 6.0


def bary4(p, a, b, c, d):  # is_mesh/util.h:349-364
    co = [signed_volume(p, b, c, d), signed_volume(a, p, c, d), signed_volume(a, b, p, d), signed_volume(a, b, c, p)]
    s = co[0] + co[1] + co[2] + co[3]
    return [x / s for x in co]


def tet_level(n):
    d = 1.0 / float(n)
    pts = [(0.0, 0.0, 0.0), (0.0, 1.0, 0.0), (1.0, 0.0, 0.0), (0.0, 0.0, 1.0)]
    rows = []
    for i in range(n):
        for j in range(n):
            for k in range(n):
                A1 = (i * d, j * d, k * d)
                A2 = add(A1, (d, 0.0, 0.0))
                A3 = add(A1, (d, d, 0.0))
                A4 = add(A1, (0.0, d, 0.0))
                A5 = add(A1, (0.0, 0.0, d))
                A6 = add(A1, (d, 0.0, d))
                A7 = add(A1, (d, d, d))
                A8 = add(A1, (0.0, d, d))
                for t in ((A1, A4, A2, A5), (A5, A4, A2, A6), (A5, A6, A4, A8), (A2, A3, A4, A6), (A3, A4, A6, A8), (A3, A6, A7, A8)):
                    s = add(add(add(t[0], t[1]), t[2]), t[3])
                    c = (s[0] / 4, s[1] / 4, s[2] / 4)
                    if c[0] + c[1] + c[2] < 1:
                        co = bary4(c, *pts)
                        rows.append([float("%f" % x) for x in co])
    return rows


def tri_level(n):
    tris = [(1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)]

    def comb(e1, e2, e3):
        a = (tris[0][0] * e1, tris[0][1] * e1, tris[0][2] * e1)
        b = (tris[1][0] * e2, tris[1][1] * e2, tris[1][2] * e2)
        c = (tris[2][0] * e3, tris[2][1] * e3, tris[2][2] * e3)
        return add(add(a, b), c)

    res = n
    step = 1.0 / float(res)
    rows = []
    for i in range(res):
        s1 = i * step
        for j in range(res - i, 0, -1):
            s2 = j * step
            ep1 = s1 + step / 3.0
            ep2 = s2 - step * 2.0 / 3.0
            ep3 = 1 - ep1 - ep2
            rows.append([float("%f" % x) for x in comb(ep1, ep2, ep3)])
            ep1 = s1 + step * 2 / 3
            ep2 = s2 - step * 4 / 3
            if ep2 > 0:
                ep3 = 1 - ep1 - ep2
                rows.append([float("%f" % x) for x in comb(ep1, ep2, ep3)])
    return rows


# Gauss quadrature sets for triangles: (weight, b0, b1, b2); numeric constants of the
# algorithm (DEMO/gaussian_points.h:24-81).  Set k is used when the face area falls in
# bracket k of GAUSS_AREA_THRES (gaussian_points.h:97-101).
GAUSS = [
    [(1, 0.333333, 0.333333, 0.333333)],
    [(0.333333, 0.166660, 0.166660, 0.666660), (0.333333, 0.166660, 0.666660, 0.166660), (0.333333, 0.666660, 0.166660, 0.166660)],
    [(-0.562500, 0.333330, 0.333330, 0.333330), (0.520833, 0.200000, 0.200000, 0.600000), (0.520833, 0.200000, 0.600000, 0.200000),
     (0.520833, 0.600000, 0.200000, 0.200000)],
    [(0.109952, 0.091570, 0.091570, 0.816840), (0.109952, 0.091570, 0.816840, 0.091570), (0.109952, 0.816840, 0.091570, 0.091570),
     (0.223382, 0.108100, 0.445940, 0.445940), (0.223382, 0.445940, 0.108100, 0.445940), (0.223382, 0.445940, 0.445940, 0.108100)],
    [(0.125939, 0.101280, 0.101280, 0.797420), (0.125939, 0.101280, 0.797420, 0.101280), (0.125939, 0.797420, 0.101280, 0.101280),
     (0.132394, 0.059710, 0.470140, 0.470140), (0.132394, 0.470140, 0.059710, 0.470140), (0.132394, 0.470140, 0.470140, 0.059710),
     (0.225000, 0.333330, 0.333330, 0.333330)],
    [(0.063691, 0.037470, 0.165400, 0.797100), (0.063691, 0.037470, 0.797100, 0.165400), (0.063691, 0.165400, 0.037470, 0.797100),
     (0.063691, 0.165400, 0.797100, 0.037470), (0.063691, 0.797100, 0.037470, 0.165400), (0.063691, 0.797100, 0.165400, 0.037470),
     (0.205951, 0.124940, 0.437520, 0.437520), (0.205951, 0.437520, 0.124940, 0.437520), (0.205951, 0.437520, 0.437520, 0.124940)],
    [(0.050840, 0.063080, 0.063080, 0.873820), (0.050840, 0.063080, 0.873820, 0.063080), (0.050840, 0.873820, 0.063080, 0.063080),
     (0.082850, 0.053140, 0.310350, 0.636500), (0.082850, 0.053140, 0.636500, 0.310350), (0.082850, 0.310350, 0.053140, 0.636500),
     (0.082850, 0.310350, 0.636500, 0.053140), (0.082850, 0.636500, 0.053140, 0.310350), (0.082850, 0.636500, 0.310350, 0.053140),
     (0.116780, 0.249280, 0.249280, 0.501420), (0.116780, 0.249280, 0.501420, 0.249280), (0.116780, 0.501420, 0.249280, 0.249280)],
]
GAUSS_AREA_THRES = [6, 10, 20, 30, 50, 80, 999999]


def build():
    tet = [tet_level(n) for n in range(1, 10)]
    tri = [tri_level(n) for n in range(1, 15)]
    return tet, tri


def checksum(tet, tri):
    h = hashlib.sha256()
    for lv in tet:
        for r in lv:
            h.update(struct.pack("<4d", *r))
    for lv in tri:
        for r in lv:
            h.update(struct.pack("<3d", *r))
    for s in GAUSS:
        for r in s:
            h.update(struct.pack("<4d", *[float(x) for x in r]))
    h.update(struct.pack("<7d", *[float(x) for x in GAUSS_AREA_THRES]))
    return h.hexdigest()


def parse_reference(ref):
    """Parse the literal tables out of the reference source (container only)."""
    txt = open(ref + "/DEMO/tet_dis_coord.cpp").read()
    a = txt.index("tet_dis_coord =")
    b = txt.index("tri_dis_coord =")
    c = txt.index("std::vector<size_t> dis_coord_size")
    num = r"-?\d+\.\d+"
    t4 = re.findall(r"\{\s*(%s),\s*(%s),\s*(%s),\s*(%s)\s*\}" % (num, num, num, num), txt[a:b])
    t3 = re.findall(r"\{\s*(%s),\s*(%s),\s*(%s)\s*\}" % (num, num, num), txt[b:c])
    g = open(ref + "/DEMO/gaussian_points.h").read()
    gp = re.findall(r"\{\s*(-?[\d.]+),\s*\{\s*(%s),\s*(%s),\s*(%s)\s*\}\s*\}" % (num, num, num), g)
    th = re.search(r"thres = \{([^}]*)\}", g).group(1)
    return ([[float(x) for x in r] for r in t4], [[float(x) for x in r] for r in t3],
            [[float(x) for x in r] for r in gp], [float(x) for x in th.split(",")])


def emit(path, tet, tri):
    out = []
    w = out.append
    w("// GENERATED by tools/gen_tables.py -- do not edit.")
    w("// Barycentric sample tables of the DSC image-energy path.  Tet levels n=1..9 hold n^3 points")
    w("// (reference: DEMO/tet_dis_coord.cpp:12-2066, generator DEMO/image3d.cpp:256-306); triangle")
    w("// levels n=1..14 hold n^2 points (DEMO/tet_dis_coord.cpp:2068-3127); Gauss sets follow")
    w("// DEMO/gaussian_points.h:24-101.  Values are 6-decimal literals on purpose: the sample")
    w("// positions, hence the voxel indices, depend on these exact doubles.")
    w("// sha256 of the packed doubles: " + checksum(tet, tri))
    w("#pragma once")
    w("#define DSC_TET_LEVELS 9")
    w("#define DSC_TRI_LEVELS 14")
    w("#define DSC_GAUSS_SETS 7")
    offs = [0]
    for lv in tet:
        offs.append(offs[-1] + len(lv))
    w("#define DSC_TET_POINTS_TOTAL %d" % offs[-1])
    w("static const int DSC_TET_LEVEL_OFFSET[DSC_TET_LEVELS + 1] = {%s};" % ", ".join(str(o) for o in offs))
    w("// DSC_TET_BARY[4*i+k]: k-th barycentric weight of flat sample point i")
    w("static const double DSC_TET_BARY[4 * DSC_TET_POINTS_TOTAL] = {")
    for n, lv in enumerate(tet):
        w("    // n = %d" % (n + 1))
        for r in lv:
            w("    %s," % ", ".join("%f" % x for x in r))
    w("};")
    toffs = [0]
    for lv in tri:
        toffs.append(toffs[-1] + len(lv))
    w("#define DSC_TRI_POINTS_TOTAL %d" % toffs[-1])
    w("static const int DSC_TRI_LEVEL_OFFSET[DSC_TRI_LEVELS + 1] = {%s};" % ", ".join(str(o) for o in toffs))
    w("static const double DSC_TRI_BARY[3 * DSC_TRI_POINTS_TOTAL] = {")
    for n, lv in enumerate(tri):
        w("    // n = %d" % (n + 1))
        for r in lv:
            w("    %s," % ", ".join("%f" % x for x in r))
    w("};")
    goffs = [0]
    for s in GAUSS:
        goffs.append(goffs[-1] + len(s))
    w("#define DSC_GAUSS_POINTS_TOTAL %d" % goffs[-1])
    w("static const int DSC_GAUSS_SET_OFFSET[DSC_GAUSS_SETS + 1] = {%s};" % ", ".join(str(o) for o in goffs))
    w("static const double DSC_GAUSS_AREA_THRES[DSC_GAUSS_SETS] = {%s};" % ", ".join(str(t) for t in GAUSS_AREA_THRES))
    w("// DSC_GAUSS[4*i+0] = weight, [4*i+1..3] = barycentric coordinates of Gauss point i")
    w("static const double DSC_GAUSS[4 * DSC_GAUSS_POINTS_TOTAL] = {")
    for s in GAUSS:
        w("    // %d point%s" % (len(s), "" if len(s) == 1 else "s"))
        for r in s:
            w("    %s," % ", ".join(("%f" % float(x)) for x in r))
    w("};")
    open(path, "w").write("\n".join(out) + "\n")


def main():
    tet, tri = build()
    if "--check" in sys.argv:
        ref = sys.argv[sys.argv.index("--check") + 1]
        t4, t3, gp, th = parse_reference(ref)
        flat4 = [r for lv in tet for r in lv]
        flat3 = [r for lv in tri for r in lv]
        flatg = [[float(x) for x in r] for s in GAUSS for r in s]
        assert len(t4) == len(flat4), (len(t4), len(flat4))
        assert t4 == flat4, "tet table differs from the reference literals"
        assert len(t3) == len(flat3), (len(t3), len(flat3))
        assert t3 == flat3, "tri table differs from the reference literals"
        assert gp == flatg, "gauss sets differ from the reference literals"
        assert th == [float(x) for x in GAUSS_AREA_THRES]
        print("tables identical to the reference literals: %d tet points, %d tri points, %d gauss points" % (len(t4), len(t3), len(gp)))
    print("sha256", checksum(tet, tri))
    if len(sys.argv) > 1 and not sys.argv[1].startswith("--"):
        emit(sys.argv[1], tet, tri)


if __name__ == "__main__":
    main()
