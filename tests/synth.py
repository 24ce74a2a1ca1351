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


def wkb_to_soa(wkbs):
    """Little-endian 2-D ISO WKB list -> the SoA form of rz_geom_soa (geom_part_off, part_kind, part_seq_off,
    seq_coord_off, x, y) with the reference's pooling rules (burn_geometry.rs:24-210): a test helper that walks the
    bytes in Python, independent of the C++ readers."""
    import struct

    gpo, kinds, pso, sco, xs, ys = [0], [], [0], [0], [], []

    def seq(buf, o):
        (n,) = struct.unpack_from("<I", buf, o)
        o += 4
        c = np.frombuffer(buf, "<f8", 2 * n, o).reshape(n, 2)
        xs.append(c[:, 0])
        ys.append(c[:, 1])
        sco.append(sco[-1] + n)
        return o + 16 * n

    def rings(buf, o):
        (nr,) = struct.unpack_from("<I", buf, o)
        o += 4
        for _ in range(nr):
            o = seq(buf, o)
        return o

    def geom(buf, o, open_part=None):
        assert buf[o] == 1
        (t,) = struct.unpack_from("<I", buf, o + 1)
        o += 5
        kind = {1: 2, 4: 2, 2: 1, 5: 1, 3: 0, 6: 0}.get(t)
        if t == 7:
            (n,) = struct.unpack_from("<I", buf, o)
            o += 4
            for _ in range(n):
                o = geom(buf, o)
            return o
        own = open_part is None
        if own:
            kinds.append(kind)
        if t == 1:
            xs.append(np.frombuffer(buf, "<f8", 1, o))
            ys.append(np.frombuffer(buf, "<f8", 1, o + 8))
            sco.append(sco[-1] + 1)
            o += 16
        elif t == 2:
            o = seq(buf, o)
        elif t == 3:
            o = rings(buf, o)
        else:
            (n,) = struct.unpack_from("<I", buf, o)
            o += 4
            for _ in range(n):
                o = geom(buf, o, open_part=kind)
        if own:
            pso.append(len(sco) - 1)
        return o

    for b in wkbs:
        geom(bytes(b), 0)
        gpo.append(len(kinds))
    cat = lambda v: np.concatenate(v) if v else np.empty(0)  # noqa: E731
    return (np.array(gpo, np.uint64), np.array(kinds, np.uint8), np.array(pso, np.uint64), np.array(sco, np.uint64),
            np.ascontiguousarray(cat(xs), np.float64), np.ascontiguousarray(cat(ys), np.float64))


def _ragged_positions(cnt):
    """(owner index, position inside the owner) of every element of a ragged array with row lengths `cnt`."""
    cnt = np.asarray(cnt, np.int64)
    start = np.cumsum(cnt) - cnt
    owner = np.repeat(np.arange(len(cnt)), cnt)
    return owner, np.arange(int(cnt.sum())) - np.repeat(start, cnt)


def config2_soa(seed=2, n=100_000, size=16384):
    """BASELINE config 2, SplitMix64-seeded: n mixed geometries on a size x size grid in input order -
    60 % star polygons (V ~ U{16..64}, rho 64), 25 % LineStrings (n ~ U{2..32} vertices, random walk, step <= 64
    px), 15 % Points / MultiPoints (1..8 points).  One part and one sequence per geometry.
    -> (geom_part_off, part_kind, part_seq_off, seq_coord_off, x, y), the rz_geom_soa arrays."""
    u = splitmix_u(seed, n, 0)
    kind = np.where(u < 0.60, 0, np.where(u < 0.85, 1, 2)).astype(np.uint8)
    ip, il, iq = (np.flatnonzero(kind == k) for k in (0, 1, 2))
    cnt = np.zeros(n, np.int64)
    # polygons: the star generator of configs 1/3/4 on the polygon subset
    px, py, poff = star_polygons(seed, len(ip), 16, 64, 64.0, size, size)
    cnt[ip] = np.diff(poff.astype(np.int64))
    # line strings: start ~ U(extent), steps ~ U(-45, 45)^2 (|step| <= 64)
    nl = (2 + np.floor(splitmix_u(seed, len(il), 6) * 31)).astype(np.int64)
    cnt[il] = nl
    own, k = _ragged_positions(nl)
    tot = len(own)
    sx = (splitmix_u(seed, tot, 7) - 0.5) * 90.0
    sy = (splitmix_u(seed, tot, 8) - 0.5) * 90.0
    sx[k == 0] = (splitmix_u(seed, len(il), 9) * size)
    sy[k == 0] = (splitmix_u(seed, len(il), 10) * size)
    cx, cy = np.cumsum(sx), np.cumsum(sy)
    first = np.cumsum(nl) - nl
    base_x = np.repeat(cx[first] - sx[first], nl)
    base_y = np.repeat(cy[first] - sy[first], nl)
    lx, ly = cx - base_x, cy - base_y
    # points: 1..8 per geometry, ~ U(extent)
    nq = (1 + np.floor(splitmix_u(seed, len(iq), 11) * 8)).astype(np.int64)
    cnt[iq] = nq
    qx = splitmix_u(seed, int(nq.sum()), 12) * size
    qy = splitmix_u(seed, int(nq.sum()), 13) * size
    # interleave the three coordinate sets in geometry order
    off = np.zeros(n + 1, np.int64)
    off[1:] = np.cumsum(cnt)
    x = np.empty(int(off[-1]))
    y = np.empty(int(off[-1]))
    for idx, c, sxs, sys_ in ((ip, cnt[ip], px, py), (il, nl, lx, ly), (iq, nq, qx, qy)):
        own, k = _ragged_positions(c)
        dst = off[idx][own] + k
        x[dst] = sxs
        y[dst] = sys_
    ar = np.arange(n + 1, dtype=np.uint64)
    return ar, kind, ar.copy(), off.astype(np.uint64), x, y


def soa_select(soa, keep):
    """The geometries with keep[i] (one part / one sequence per geometry, as config2_soa builds them), order
    preserved -> the same six arrays."""
    gpo, kind, pso, sco, x, y = soa
    o = sco.astype(np.int64)
    cnt = (o[1:] - o[:-1])[keep]
    own, k = _ragged_positions(cnt)
    src = o[:-1][keep][own] + k
    noff = np.zeros(len(cnt) + 1, np.uint64)
    noff[1:] = np.cumsum(cnt)
    ar = np.arange(len(cnt) + 1, dtype=np.uint64)
    return ar, kind[keep], ar.copy(), noff, x[src], y[src]


def soa_to_wkb(soa):
    """One-part / one-sequence SoA geometries -> WKB list (for the oracle, which reads WKB)."""
    from oracle import wkt2wkb as W

    _, kind, _, sco, x, y = soa
    out = []
    for i in range(len(kind)):
        a, b = int(sco[i]), int(sco[i + 1])
        p = np.stack([x[a:b], y[a:b]], 1)
        if kind[i] == 0:
            out.append(W.polygon_wkb([p]))
        elif kind[i] == 1:
            out.append(W.linestring_wkb(p))
        else:
            out.append(W.point_wkb(*p[0]) if len(p) == 1 else W.multipoint_wkb(p))
    return out


def config3(seed=3, n=100_000, size=8192):
    """BASELINE config 3: n star polygons (V = 64, rho 256) on size^2, by = str(i mod 32), int32 field."""
    x, y, off = star_polygons(seed, n, 64, 64, 256.0, size, size)
    idx = np.arange(n, dtype=np.int64)
    field = (1 + (idx * 2654435761 % 10**6)).astype(np.int32)
    by = [str(i % 32) for i in range(n)]
    return x, y, off, field, by
