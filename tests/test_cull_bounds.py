"""CPU: the geometric claims behind the raster's visibility shortcuts.

`cube_hidden` drops a cube, and `triangle_hidden` a triangle, when every pixel it could touch already holds a nearer
depth.  Both decide in fp32 from bounds: a pixel range that must contain every fragment the reference would produce
(render.c:94-115: a fragment exists where a pixel CENTRE passes the inside test) and a depth cap no fragment can exceed.
If a bound were too tight for some camera, a visible triangle would silently disappear.  On the GPU this is checked by
rendering with and without the shortcuts (tests/test_gpu_parity.py, tools/gpu_shortcut_check.py) -- with the demo camera.
Here the bounds are restated in numpy float32 and checked against the reference's own fragments (a float64 restatement of
render.c:48-163 on the oracle's triangles) for random cameras, coordinates and screens (tests/fuzz_cases.py): every
fragment of a triangle lies inside the pixel range of its cube's test and of its own test, at or below their depth caps.
The fp32 coverage pre-test of `raster_triangle` (a pixel is dropped before the exact fp64 inside test when an edge function
is below minus its error bound) is restated too: it must keep every pixel the reference covers.

Triangles the shortcuts leave alone by construction are skipped here as there: collapsed ones (a vertex within rounding
of a cube corner: `patch_irregular` / the `fr` guard), vertices near the screen plane, bounds that are not finite.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from fuzz_cases import random_case  # noqa: E402
from oracle import oracle  # noqa: E402

F = np.float32
# cube-edge id (oracle/c3d_oracle.c: cube_edge) -> the two cube corners it joins (corner bit0 = +x, bit1 = +y, bit2 = +z)
EDGE_ENDS = np.array(
    [(0, 1), (2, 3), (4, 5), (6, 7), (0, 2), (1, 3), (4, 6), (5, 7), (0, 4), (1, 5), (2, 6), (3, 7),
     (0, 6), (1, 7), (0, 5), (2, 7), (0, 3), (4, 7), (0, 7)], dtype=np.int64)


def fma32(a, b, c):
    """fp32 fused multiply-add: the product of two floats is exact in double"""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


class Camera:
    def __init__(self, scene):
        cam, center = np.array(scene["camera"], dtype=np.float64), np.array(scene["center"], dtype=np.float64)
        lx, ly = np.array(scene["local_x"], dtype=np.float64), np.array(scene["local_y"], dtype=np.float64)
        n = np.cross(lx, ly)
        n = n * (1.0 / np.sqrt(n @ n))
        self.cam, self.center, self.lx, self.ly, self.n = cam, center, lx, ly, n
        self.w, self.h = int(scene["width"]), int(scene["height"])
        self.W, self.H = float(self.w), float(self.h)
        # make_cull(), c3d_runtime.cu: constants rounded to fp32 once
        lxlx, lyly = lx @ lx, ly @ ly
        c = n @ center - n @ cam
        self.q_cam, self.q_n = cam.astype(F), n.astype(F)
        self.q_lxs, self.q_lys = (lx / lxlx).astype(F), (ly / lyly).astype(F)
        self.q_ax, self.q_ay = F(((cam - center) @ lx) / lxlx), F(((cam - center) @ ly) / lyly)
        self.q_c = F(c)
        self.q_inv_abs_c = F(1.0 / abs(c))
        self.q_lx_norm = F(np.sqrt(lxlx) / lxlx * 1.000001)
        self.q_ly_norm = F(np.sqrt(lyly) / lyly * 1.000001)
        self.q_W, self.q_H = F(self.W), F(self.H)


def reference_fragments(cm, tri_v):
    """render.c:48-163 in float64 for one triangle (3, 3): pixel columns, rows and depths of its fragments, or None when the
    reference drops the triangle (a vertex fails project.c:28-65); also whether the clamped area was used, the projected
    vertices and the inverse area"""
    q = np.zeros((3, 3))
    for m in range(3):
        ray = tri_v[m] - cm.cam
        nr = cm.n @ ray
        if -1e-8 < nr < 1e-8:
            return None
        t = (cm.n @ cm.center - cm.n @ cm.cam) / nr
        if t <= 0.0 or 1.0 <= t:
            return None
        delta = cm.cam + t * ray - cm.center
        q[m] = ((delta @ cm.lx) / (cm.lx @ cm.lx), (delta @ cm.ly) / (cm.ly @ cm.ly), nr)
    clamp = lambda lo, hi, v: min(hi, max(lo, v))
    xmin = int((0.5 + q[:, 0].min()) * cm.W - 1)
    xmax = int((0.5 + q[:, 0].max()) * cm.W + 1)
    ymin = int((0.5 + q[:, 1].min()) * cm.H - 1)
    ymax = int((0.5 + q[:, 1].max()) * cm.H + 1)
    imin, imax = clamp(0, cm.w - 1, xmin), clamp(0, cm.w - 1, xmax)
    jmin, jmax = clamp(0, cm.h - 1, ymin), clamp(0, cm.h - 1, ymax)
    edge = lambda a, b, px, py: (px - a[0]) * (b[1] - a[1]) - (py - a[1]) * (b[0] - a[0])
    area = edge(q[0], q[1], q[2][0], q[2][1])
    clamped = abs(area) < 1e-16
    area_inv = 1.0 / (min(area, -1e-16) if area < 0.0 else max(area, 1e-16))
    ii, jj = np.meshgrid(np.arange(imin, imax + 1), np.arange(jmin, jmax + 1))
    px, py = (ii + 0.5) / cm.W - 0.5, (jj + 0.5) / cm.H - 0.5
    w0, w1, w2 = area_inv * edge(q[1], q[2], px, py), area_inv * edge(q[2], q[0], px, py), area_inv * edge(q[0], q[1], px, py)
    inside = ~((w0 < -1e-8) | (w1 < -1e-8) | (w2 < -1e-8))
    with np.errstate(divide="ignore", invalid="ignore"):
        dist = 1.0 / (w0 / q[0, 2] + w1 / q[1, 2] + w2 / q[2, 2])
    return ii[inside], jj[inside], dist[inside], clamped, q, area_inv


def coverage_pretest_keeps(cm, q, area_inv, ii, jj):
    """raster_triangle(), c3d_kernels.cu: before the exact fp64 inside test a pixel is dropped when one of the three edge
    functions, evaluated in fp32, lies below -tol (tol = a bound of the fp32 evaluation error + the reference's own 1e-8
    slack).  True per pixel = not dropped."""
    sgn = -1.0 if area_inv < 0.0 else 1.0
    slack = 1.0001e-8 / abs(area_inv)
    px = ((ii + 0.5) / cm.W - 0.5).astype(F)          # k_pixel_tables: the fp64 pixel centre, rounded once
    py = ((jj + 0.5) / cm.H - 0.5).astype(F)
    keep = np.ones(len(ii), dtype=bool)
    for m in range(3):
        a, b = q[(m + 1) % 3], q[(m + 2) % 3]
        ey, ex = b[1] - a[1], b[0] - a[0]
        fax, fay, fex, fey = F(a[0]), F(a[1]), F(sgn * ex), F(sgn * ey)
        tol = 4e-7 * ((0.5 + abs(a[0])) * abs(ey) + (0.5 + abs(a[1])) * abs(ex)) + slack
        ntol = -(F(F(tol) * F(1.000001)) + F(1e-37))
        r = ((py - fay) * fex).astype(F)
        e = fma32((px - fax).astype(F), np.full(len(ii), fey, dtype=F), -r)
        keep &= ~(e < ntol)
    return keep


def project32(cm, rx, ry, rz, fused):
    """the fp32 projection shared by cube_hidden (separate roundings) and triangle_hidden (fused)"""
    n, lxs, lys = cm.q_n, cm.q_lxs, cm.q_lys
    if fused:
        nr = fma32(n[2], rz, fma32(n[1], ry, n[0] * rx))
        rl = fma32(rz, lxs[2], fma32(ry, lxs[1], rx * lxs[0]))
        rm = fma32(rz, lys[2], fma32(ry, lys[1], rx * lys[0]))
    else:
        nr = (n[0] * rx + n[1] * ry) + n[2] * rz
        rl = (lxs[0] * rx + lxs[1] * ry) + lxs[2] * rz
        rm = (lys[0] * rx + lys[1] * ry) + lys[2] * rz
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        t = (cm.q_c / nr).astype(F)
    return nr, t, rl, rm


def cube_tests(cm, grids, ci, cj, ck):
    """cube_hidden(), c3d_kernels.cu, per cube: (decides, imin, imax, jmin, jmax, zcap)"""
    gf = [g.astype(F) for g in grids]
    big = F(3e38)
    xlo = np.full(len(ci), big); xhi = -xlo; ylo = xlo.copy(); yhi = xhi.copy(); tlo = xlo.copy(); thi = xhi.copy(); znear = xhi.copy()
    with np.errstate(invalid="ignore", over="ignore"):
        for corner in range(8):
            a, b, d = corner & 1, (corner >> 1) & 1, corner >> 2
            rx, ry, rz = gf[0][ci + a] - cm.q_cam[0], gf[1][cj + b] - cm.q_cam[1], gf[2][ck + d] - cm.q_cam[2]
            nr, t, rl, rm = project32(cm, rx, ry, rz, fused=False)
            sx, sy = (cm.q_ax + t * rl).astype(F), (cm.q_ay + t * rm).astype(F)
            xlo, xhi = np.fmin(xlo, sx), np.fmax(xhi, sx)
            ylo, yhi = np.fmin(ylo, sy), np.fmax(yhi, sy)
            tlo, thi = np.fmin(tlo, t), np.fmax(thi, t)
            znear = np.fmax(znear, nr)
        decides = (tlo > F(1e-4)) & (thi < F(0.9999)) & (xlo > F(-1e6)) & (xhi < F(1e6)) & (ylo > F(-1e6)) & (yhi < F(1e6))
        X0, X1 = (F(0.5) + xlo) * cm.q_W - F(0.55), (F(0.5) + xhi) * cm.q_W - F(0.45)
        Y0, Y1 = (F(0.5) + ylo) * cm.q_H - F(0.55), (F(0.5) + yhi) * cm.q_H - F(0.45)
        safe = lambda v: np.nan_to_num(v, nan=0.0, posinf=1e9, neginf=-1e9)
        imin, imax = np.maximum(0, np.ceil(safe(X0))), np.minimum(cm.w - 1, np.floor(safe(X1)))
        jmin, jmax = np.maximum(0, np.ceil(safe(Y0))), np.minimum(cm.h - 1, np.floor(safe(Y1)))
    zn = znear.astype(np.float64)
    zcap = zn + (1e-4 * np.abs(zn) + 1e-6)
    return decides, imin, imax, jmin, jmax, zcap


def triangle_test(cm, grids, field, thr, ci, cj, ck, edges):
    """triangle_hidden(), c3d_kernels.cu, for one triangle: None when it leaves the triangle to the exact path, else
    (imin, imax, jmin, jmax, zcap) with imax < imin when it claims that no pixel centre can be covered"""
    gf = [g.astype(F) for g in grids]
    thr32 = F(thr)
    xlo = ylo = F(3e38); xhi = yhi = znear = F(-3e38); zslack = F(0)
    for m in range(3):
        tail, head = EDGE_ENDS[edges[m]]
        corner = lambda c: field[ck + (c >> 2), cj + ((c >> 1) & 1), ci + (c & 1)]
        v0, v1 = F(corner(tail)), F(corner(head))
        dv = F(v1 - v0)
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            inv = F(F(1) / dv)
            ferr = F(F(F(F(2.00002e-7) * F(F(abs(thr32) + abs(v0)) + abs(v1))) * abs(inv)) + F(3e-7))
        if not (ferr < F(0.02)) or not (abs(dv) > F(1e-7)):
            return None
        fr = min(max(F(F(thr32 - v0) * inv), F(0)), F(1))
        if not (F(fr - ferr) > F(1e-5) and F(fr + ferr) < F(0.99999)):
            return None
        pos = lambda c: (gf[0][ci + (c & 1)], gf[1][cj + ((c >> 1) & 1)], gf[2][ck + (c >> 2)])
        (x0, y0, z0), (x1, y1, z1) = pos(tail), pos(head)
        ex, ey, ez = F(x1 - x0), F(y1 - y0), F(z1 - z0)
        one = lambda v: np.array([v], dtype=F)
        rx = fma32(one(fr), one(ex), one(x0)) - cm.q_cam[0]
        ry = fma32(one(fr), one(ey), one(y0)) - cm.q_cam[1]
        rz = fma32(one(fr), one(ez), one(z0)) - cm.q_cam[2]
        nr, t, rl, rm = project32(cm, rx, ry, rz, fused=True)
        if not (t[0] > F(1e-4) and t[0] < F(0.9999)):
            return None
        X = fma32(fma32(t, rl, one(F(0.5) + cm.q_ax)), one(cm.q_W), one(F(-0.5)))[0]
        Y = fma32(fma32(t, rm, one(F(0.5) + cm.q_ay)), one(cm.q_H), one(F(-0.5)))[0]
        span = F(ferr * F(F(abs(ex) + abs(ey)) + abs(ez)))
        tc = F(t[0] * cm.q_inv_abs_c)
        st = F(F(F(1.5) * span) * t[0])
        mx = fma32(one(F(st * cm.q_W)), fma32(one(abs(rl[0])), one(tc), one(cm.q_lx_norm)), one(F(0.05)))[0]
        my = fma32(one(F(st * cm.q_H)), fma32(one(abs(rm[0])), one(tc), one(cm.q_ly_norm)), one(F(0.05)))[0]
        xlo, xhi = min(xlo, F(X - mx)), max(xhi, F(X + mx))
        ylo, yhi = min(ylo, F(Y - my)), max(yhi, F(Y + my))
        znear = max(znear, nr[0])
        zslack = max(zslack, span)
    if not (xlo > F(-1e6) and xhi < F(1e6) and ylo > F(-1e6) and yhi < F(1e6)):
        return None
    imin, imax = max(0, int(np.ceil(xlo))), min(cm.w - 1, int(np.floor(xhi)))
    jmin, jmax = max(0, int(np.ceil(ylo))), min(cm.h - 1, int(np.floor(yhi)))
    zn = float(znear)
    zcap = zn + (1e-4 * abs(zn) + 1e-6 + 2.0 * float(zslack))
    if imax < imin or jmax < jmin:
        return (0, -1, 0, -1, zcap)
    return (imin, imax, jmin, jmax, zcap)


def cubes_regular(field, thr, ci, cj, ck):
    """patch_irregular(), c3d_sweep.cu, applied to the eight corners of each cube (the kernel applies it to the whole 4 x 8 patch
    the cube lies in and so marks MORE cubes irregular; what is checked here is therefore a superset of what the raster may
    cull): no corner value within 1e-6 (relative to the largest distance) or 1e-8 (absolute) of the iso-value"""
    a = np.stack([np.abs(field[ck + (c >> 2), cj + ((c >> 1) & 1), ci + (c & 1)].astype(np.float64) - thr) for c in range(8)])
    amin, amax = a.min(axis=0), a.max(axis=0)
    return np.isfinite(a).all(axis=0) & (amin > 1.1e-6 * amax) & (amin > 1.01e-8)


def check_case(seed, flip=False, max_triangles=700):
    scene, grids, f, contours, _, _, _ = random_case(seed)
    if flip:
        # local_x x local_y pointing away from the camera: the reference then keeps the FARTHEST surface (render.c:125-132
        # compares the signed n . ray); the bounds must hold for that orientation too
        scene = dict(scene, local_y=tuple(-v for v in scene["local_y"]))
    cm = Camera(scene)
    nz, ny, nx = f.shape
    rng = np.random.default_rng(seed)
    stats = dict(triangles=0, fragments=0, cube_claims=0, tri_claims=0, tri_empty_claims=0)
    for c in contours:
        tris, _ = oracle.tessellate(f, grids, c["threshold"])
        if 0 == len(tris):
            continue
        cube = tris["cube"].astype(np.int64)
        ci, cj, ck = cube % (nx - 1), (cube // (nx - 1)) % (ny - 1), cube // ((nx - 1) * (ny - 1))
        decides, imin, imax, jmin, jmax, zcap = cube_tests(cm, grids, ci, cj, ck)
        regular = cubes_regular(f, c["threshold"], ci, cj, ck)
        pick = np.arange(len(tris))
        if len(pick) > max_triangles:
            pick = np.sort(rng.choice(pick, max_triangles, replace=False))
        for n in pick:
            got = reference_fragments(cm, tris["v"][n])
            if got is None:
                continue
            ii, jj, dist, clamped, q, area_inv = got
            kept = coverage_pretest_keeps(cm, q, area_inv, ii, jj)
            assert kept.all(), f"seed {seed}: triangle {n}: the fp32 coverage pre-test drops {int((~kept).sum())} of {len(ii)} covered pixels"
            stats["triangles"] += 1
            stats["fragments"] += len(ii)
            own = triangle_test(cm, grids, f, c["threshold"], ci[n], cj[n], ck[n], tris["edge"][n])
            if own is not None:
                # not collapsed, not near the screen plane: the triangle's own box and cap must hold
                stats["tri_claims"] += 1
                stats["tri_empty_claims"] += own[1] < own[0]
                assert not clamped
                ok = (ii >= own[0]) & (ii <= own[1]) & (jj >= own[2]) & (jj <= own[3]) & (dist <= own[4])
                assert ok.all(), f"seed {seed}: triangle {n}: {int((~ok).sum())} of {len(ii)} fragments outside triangle_hidden's bounds"
            if decides[n] and regular[n]:
                # the cube's claim covers every triangle of a cube that k_expand has not marked irregular
                stats["cube_claims"] += 1
                ok = (ii >= imin[n]) & (ii <= imax[n]) & (jj >= jmin[n]) & (jj <= jmax[n]) & (dist <= zcap[n])
                assert ok.all(), f"seed {seed}: triangle {n}: {int((~ok).sum())} of {len(ii)} fragments outside cube_hidden's bounds"
    return stats


@pytest.mark.parametrize("seed", range(1000, 1030))
def test_fragments_lie_inside_the_bounds_of_the_visibility_tests(seed):
    check_case(seed)


@pytest.mark.parametrize("seed", range(1030, 1038))
def test_bounds_hold_when_the_screen_normal_points_away_from_the_camera(seed):
    check_case(seed, flip=True)


def test_the_check_has_teeth():
    """the same cases must produce plenty of fragments, of claims by both tests, and of 'no pixel centre' claims"""
    total = dict(triangles=0, fragments=0, cube_claims=0, tri_claims=0, tri_empty_claims=0)
    for seed in (1004, 1009, 1010, 1012):
        for k, v in check_case(seed).items():
            total[k] += int(v)
    assert total["fragments"] > 20000 and total["tri_claims"] > 2000 and total["cube_claims"] > 2000, total
    assert total["tri_empty_claims"] > 50, total
