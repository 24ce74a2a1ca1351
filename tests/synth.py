"""Deterministic synthetic workloads: SplitMix64-seeded star polygons, random-walk
lines and points.  Pure numpy; shared by tests and bench.py."""
import numpy as np

_GAMMA = np.uint64(0x9E3779B97F4A7C15)


def splitmix_u(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """n uniform doubles in [0,1): u_i = (mix(seed + (i+1)*GAMMA) >> 11) * 2^-53."""
    with np.errstate(over="ignore"):
        i = np.arange(1, n + 1, dtype=np.uint64) + np.uint64(stream) * np.uint64(1 << 40)
        z = np.uint64(seed) + i * _GAMMA
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def star_polygons(seed, n_polys, vmin, vmax, rho, width, height, chunk=1 << 16):
    """-> x, y (closed rings, float64), ring_off (uint64[n_polys+1])."""
    nv = (vmin + np.floor(splitmix_u(seed, n_polys, 1) * (vmax - vmin + 1))).astype(np.int64)
    cx = splitmix_u(seed, n_polys, 2) * width
    cy = splitmix_u(seed, n_polys, 3) * height
    off = np.zeros(n_polys + 1, np.uint64)
    off[1:] = np.cumsum(nv + 1)
    total = int(off[-1])
    x = np.empty(total)
    y = np.empty(total)
    for a in range(0, n_polys, chunk):
        b = min(a + chunk, n_polys)
        cnt = nv[a:b]
        tot = int(cnt.sum())
        pid = np.repeat(np.arange(b - a), cnt)
        start = np.cumsum(cnt) - cnt
        k = np.arange(tot) - np.repeat(start, cnt)
        gk = int(off[a]) - a  # global vertex index base (without closing vertices)
        u1 = splitmix_u(seed, tot, 4)[...] if False else splitmix_u(seed + 17 * (a + 1), tot, 4)
        u2 = splitmix_u(seed + 31 * (a + 1), tot, 5)
        theta = 2.0 * np.pi * (k + 0.8 * u1) / np.repeat(cnt, cnt)
        r = rho * (0.5 + 0.5 * u2)
        px = cx[a:b][pid] + r * np.cos(theta)
        py = cy[a:b][pid] + r * np.sin(theta)
        # scatter into rings with one closing vertex each
        dst = (np.repeat(off[a:b].astype(np.int64), cnt) + k)
        x[dst] = px
        y[dst] = py
        last = off[a + 1:b + 1].astype(np.int64) - 1
        x[last] = x[off[a:b].astype(np.int64)]
        y[last] = y[off[a:b].astype(np.int64)]
        del gk
    return x, y, off


def polygons_to_wkb(x, y, off):
    from oracle.wkt2wkb import polygon_wkb

    return [polygon_wkb([np.stack([x[int(a):int(b)], y[int(a):int(b)]], 1)]) for a, b in zip(off[:-1], off[1:])]


def mixed_geometries(seed, n, width, height, rho=24.0):
    """WKB list: 60% star polygons (some with a hole / multipolygons), 25% (multi)linestrings,
    15% (multi)points, plus a few collections.  Small helper for parity tests (python loops)."""
    from oracle import wkt2wkb as W

    rng = np.random.default_rng(seed)
    out = []

    def star(cx, cy, r, nv):
        th = 2 * np.pi * (np.arange(nv) + 0.8 * rng.random(nv)) / nv
        rr = r * (0.5 + 0.5 * rng.random(nv))
        p = np.stack([cx + rr * np.cos(th), cy + rr * np.sin(th)], 1)
        return np.vstack([p, p[:1]])

    def poly():
        cx, cy = rng.random() * width, rng.random() * height
        rings = [star(cx, cy, rho, rng.integers(3, 40))]
        if rng.random() < 0.3:
            rings.append(star(cx, cy, rho * 0.3, rng.integers(3, 12)))
        return rings

    def line():
        nv = rng.integers(2, 20)
        p = np.cumsum(rng.normal(0, rho / 2, (nv, 2)), 0) + [rng.random() * width, rng.random() * height]
        if rng.random() < 0.2:
            p = np.vstack([p, p[:1]])  # closed
        if rng.random() < 0.3:
            p = np.round(p)  # lattice vertices: exercise ties
        return p

    def pts():
        k = rng.integers(1, 6)
        p = rng.random((k, 2)) * [width * 1.2, height * 1.2] - [width * 0.1, height * 0.1]
        if rng.random() < 0.3:
            p = np.vstack([p, p[:1]])  # duplicate point
        return p

    for _ in range(n):
        u = rng.random()
        if u < 0.45:
            out.append(W.polygon_wkb(poly()))
        elif u < 0.58:
            out.append(W.multipolygon_wkb([poly() for _ in range(rng.integers(1, 4))]))
        elif u < 0.72:
            out.append(W.linestring_wkb(line()))
        elif u < 0.82:
            out.append(W.multilinestring_wkb([line() for _ in range(rng.integers(1, 4))]))
        elif u < 0.95:
            p = pts()
            out.append(W.point_wkb(*p[0]) if len(p) == 1 else W.multipoint_wkb(p))
        else:
            out.append(W.collection_wkb([W.point_wkb(*pts()[0]), W.polygon_wkb(poly()), W.linestring_wkb(line()),
                                         W.collection_wkb([W.polygon_wkb(poly())])]))
    return out


def parcels(seed, n, width, height, smin=6.0, smax=14.0):
    """BASELINE config 5's geometry: n axis-jittered quads ("small parcels"), side ~U[smin, smax] px, closed rings
    of 5 vertices -> (x, y, ring_off) like star_polygons."""
    u = lambda s: splitmix_u(seed, n, s)  # noqa: E731
    cx, cy = u(0) * width, u(1) * height
    hw, hh = 0.5 * (smin + (smax - smin) * u(2)), 0.5 * (smin + (smax - smin) * u(3))
    x = np.empty((n, 5))
    y = np.empty((n, 5))
    for k, (sx, sy) in enumerate([(-1, -1), (1, -1), (1, 1), (-1, 1)]):
        x[:, k] = cx + sx * hw + (u(4 + 2 * k) - 0.5) * 2.0  # every corner jittered by up to a pixel
        y[:, k] = cy + sy * hh + (u(5 + 2 * k) - 0.5) * 2.0
    x[:, 4], y[:, 4] = x[:, 0], y[:, 0]
    off = (np.arange(n + 1, dtype=np.uint64) * 5)
    return np.ascontiguousarray(x.reshape(-1)), np.ascontiguousarray(y.reshape(-1)), off
