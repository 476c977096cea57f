"""Host-side polygon helpers behind `BaseMapping.outline / centroid / maskedByPolygon`.

Same names and argument meaning as the reference's `auromat/utils.py` (outline :136-146,
polygonArea :148-167, polygonCentroid :169-213, withoutConsecutiveDuplicates :222-232,
convexHull :234-268, pointsInsidePolygon :58-74).  The reference delegates the contour walk to
scikit-image, the hull to scipy's Delaunay and the inside test to matplotlib; none of them is
used here.  These run once per mapping on a few thousand outline nodes -- they are not on the
per-pixel path (the per-pixel inside test of `maskedByPolygon` is `amt_polygon_center_mask`).
"""
from __future__ import annotations

import numpy as np

# direction codes of a crack (the side of a defined node that faces an undefined one)
_N, _E, _S, _W = 0, 1, 2, 3
_DY = (-1, 0, 1, 0)
_DX = (0, 1, 0, -1)


def withoutConsecutiveDuplicates(arr):
    """Copy of `arr` without consecutive duplicate rows."""
    arr = np.asarray(arr)
    if len(arr) == 0:
        return arr
    keep = np.ones(len(arr), bool)
    keep[1:] = np.any(arr[1:] != arr[:-1], axis=tuple(range(1, arr.ndim))) if arr.ndim > 1 else arr[1:] != arr[:-1]
    return arr[keep]


def polygonArea(poly, signed=False):
    """Area of an unclosed polygon (n,2), shoelace formula."""
    p = np.asarray(poly, dtype=np.float64)
    x0, y0 = p[:, 0], p[:, 1]
    x1, y1 = np.roll(x0, -1), np.roll(y0, -1)
    area = 0.5 * float(np.sum(x0 * y1 - x1 * y0))
    return area if signed else abs(area)


def polygonCentroid(poly):
    """Centroid (x, y) of an unclosed polygon (n,2).  The moments are taken relative to the
    first vertex: the cross products of geographic outlines (|lon| ~ 100, node spacing ~ 1e-3)
    otherwise cancel to a few digits."""
    p = np.asarray(poly, dtype=np.float64)
    o = p[0]
    q = p - o
    x0, y0 = q[:, 0], q[:, 1]
    x1, y1 = np.roll(x0, -1), np.roll(y0, -1)
    cross = x0 * y1 - x1 * y0
    a6 = 3.0 * np.sum(cross)
    return (float(np.sum((x0 + x1) * cross) / a6 + o[0]), float(np.sum((y0 + y1) * cross) / a6 + o[1]))


def _traceContours(im):
    """All closed boundary walks of the 4-connected True regions of `im` (holes included).

    Crack following: a crack is (node, side) with the node defined and its neighbour on that
    side undefined.  Walking with the defined region on the right-hand side, from crack
    (p, d) with t = d turned clockwise:
      q = p + t undefined            -> stay on p, crack (p, t)              [convex corner]
      q and r = q + d both defined   -> crack (r, -t)                         [concave corner]
      otherwise                      -> crack (q, d)                          [straight]
    The emitted node sequence (consecutive repeats dropped) is the one a marching-squares
    iso-contour at a level just below 1 visits when its vertices are rounded to the nearest
    node, with the undefined background treated as 8-connected (reference utils.py:97-134).
    Returns a list of (n,2) int arrays in (x, y) order, clockwise in image coordinates
    (x to the right, y downwards) around defined regions."""
    im = np.asarray(im, dtype=bool)
    h, w = im.shape
    P = np.zeros((h + 2, w + 2), bool)
    P[1:-1, 1:-1] = im
    north = P[1:-1, 1:-1] & ~P[:-2, 1:-1]        # defined nodes whose upper neighbour is undefined
    starts = np.argwhere(north)                  # row-major: the first start of a walk is its top-left node
    todo = set(map(tuple, starts.tolist()))
    rows = P.tolist()                            # list-of-lists indexing is much faster than ndarray scalars
    contours = []
    for sy, sx in starts.tolist():
        if (sy, sx) not in todo:
            continue
        y, x, d = sy + 1, sx + 1, _N             # padded coordinates
        pts = []
        while True:
            if d == _N:
                todo.discard((y - 1, x - 1))
            if not pts or pts[-1] != (x, y):
                pts.append((x, y))
            t = (d + 1) & 3
            qy, qx = y + _DY[t], x + _DX[t]
            if not rows[qy][qx]:
                d = t
            else:
                ry, rx = qy + _DY[d], qx + _DX[d]
                if rows[ry][rx]:
                    y, x, d = ry, rx, (t + 2) & 3
                else:
                    y, x = qy, qx
            if y == sy + 1 and x == sx + 1 and d == _N:
                break
        if len(pts) > 1 and pts[-1] == pts[0]:
            pts.pop()
        contours.append(np.asarray(pts, dtype=np.int64) - 1)
    return contours


def outline(im):
    """Outline of a binary image whose inner structure is filled with True: the ordered
    (clockwise) nodes of its boundary as an (n,2) array in x,y order, usable as a polygon;
    concave shapes are handled.  With several regions the one of the largest area is returned
    (reference utils.py:120-132)."""
    contours = _traceContours(im)
    if not contours:
        raise ValueError('the binary image is empty')
    if len(contours) == 1:
        return contours[0]
    contours = [c for c in contours if len(c) > 2] or contours
    areas = [polygonArea(c) for c in contours]
    return contours[int(np.argmax(areas))]


def convexHull(points):
    """Convex hull of (n,2) points; vertices ordered by the angle atan2(dx, dy) about their
    mean as in the reference (utils.py:262-266).  Andrew's monotone chain; collinear points on
    hull edges are not vertices."""
    pts = np.asarray(points)
    assert pts.ndim == 2 and pts.shape[1] == 2
    u = np.unique(pts, axis=0)
    if len(u) <= 2:
        return u
    P = u.tolist()                                # sorted lexicographically by np.unique

    def half(seq):
        out = []
        for p in seq:
            while len(out) >= 2 and ((out[-1][0] - out[-2][0]) * (p[1] - out[-2][1])
                                     - (out[-1][1] - out[-2][1]) * (p[0] - out[-2][0])) <= 0:
                out.pop()
            out.append(p)
        return out

    lower, upper = half(P), half(reversed(P))
    v = np.asarray(lower[:-1] + upper[:-1], dtype=pts.dtype)
    c = v - v.mean(axis=0)
    return v[np.argsort(np.arctan2(c[:, 0], c[:, 1]), kind='stable')]


def pointsInsidePolygon(points, polygon):
    """For each point (n,2) whether it lies inside the unclosed polygon (m,2): crossing-number
    test of a ray towards +x against every edge (edges are half-open in y, so a vertex is
    never counted twice).  Same arithmetic as the device kernel `k_polygon_corner_flags`."""
    pts = np.asarray(points, dtype=np.float64)
    poly = np.asarray(polygon, dtype=np.float64)
    tx, ty = pts[:, 0], pts[:, 1]
    inside = np.zeros(len(pts), bool)
    x0, y0 = poly[-1]
    for x1, y1 in poly:
        f0, f1 = y0 >= ty, y1 >= ty
        with np.errstate(invalid='ignore'):
            hit = ((y1 - ty) * (x0 - x1) >= (x1 - tx) * (y0 - y1)) == f1
        inside ^= (f0 != f1) & hit
        x0, y0 = x1, y1
    return inside & ~(np.isnan(tx) | np.isnan(ty))


def findNearest(a, x):
    """Index of the item of the sorted sequence `a` that is closest to `x`; the left one on a tie
    (reference utils.py:270-285)."""
    import bisect
    i = bisect.bisect_left(a, x)
    if i == len(a):
        return i - 1
    if i == 0 or a[i] == x:
        return i
    return i - 1 if abs(x - a[i - 1]) <= abs(x - a[i]) else i
