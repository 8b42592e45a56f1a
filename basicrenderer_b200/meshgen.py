"""Deterministic synthetic meshes for the CLod build benchmarks and parity tests.

Vertex layout follows the reference's interleaved layout (BasicRenderer/include/Mesh/VertexLayout.h:7-36):
position f32x3 @0, normal f32x3 @12, [uv f32x2 @24]. Indices are u32 triangle lists.
All generators are pure numpy, seeded, and produce byte-identical output on every run.
"""
from __future__ import annotations

import os

import numpy as np

VERTEX_NORMALS = 1 << 1  # BasicRenderer/include/Mesh/VertexFlags.h (VERTEX_COLORS = 1 << 0)
VERTEX_TEXCOORDS = 1 << 2


class Mesh:
    """Interleaved vertex buffer + u32 index buffer."""

    def __init__(self, vertices: np.ndarray, indices: np.ndarray, flags: int, name: str):
        assert vertices.dtype == np.float32 and vertices.ndim == 2
        assert indices.dtype == np.uint32 and indices.ndim == 1 and indices.size % 3 == 0
        self.vertices = np.ascontiguousarray(vertices)
        self.indices = np.ascontiguousarray(indices)
        self.flags = flags
        self.name = name

    @property
    def vertex_count(self) -> int:
        return self.vertices.shape[0]

    @property
    def triangle_count(self) -> int:
        return self.indices.size // 3

    @property
    def stride(self) -> int:
        return self.vertices.shape[1] * 4

    @property
    def positions(self) -> np.ndarray:
        return self.vertices[:, 0:3]

    @property
    def normals(self) -> np.ndarray:
        return self.vertices[:, 3:6]


def _hash_u32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def _lattice(ix, iy, iz, seed):
    h = _hash_u32(ix.astype(np.uint32) * np.uint32(0x9E3779B1) ^ _hash_u32(iy.astype(np.uint32) * np.uint32(0x85EBCA77) ^ _hash_u32(iz.astype(np.uint32) * np.uint32(0xC2B2AE3D) ^ np.uint32(seed & 0xFFFFFFFF))))
    return (h >> np.uint32(8)).astype(np.float64) * (1.0 / 16777216.0) * 2.0 - 1.0


def value_noise3(p: np.ndarray, seed: int) -> np.ndarray:
    """Smooth value noise in [-1, 1] at float64 points p[N,3]."""
    pf = np.floor(p)
    t = p - pf
    t = t * t * (3.0 - 2.0 * t)
    i = pf.astype(np.int64)
    ix, iy, iz = i[:, 0], i[:, 1], i[:, 2]
    out = np.zeros(p.shape[0], dtype=np.float64)
    for dx in (0, 1):
        wx = t[:, 0] if dx else 1.0 - t[:, 0]
        for dy in (0, 1):
            wy = t[:, 1] if dy else 1.0 - t[:, 1]
            for dz in (0, 1):
                wz = t[:, 2] if dz else 1.0 - t[:, 2]
                out += wx * wy * wz * _lattice(ix + dx, iy + dy, iz + dz, seed)
    return out


def grid(n: int, seed: int = 1234, amplitude: float = 0.1, chunk: int = 1 << 22) -> Mesh:
    """(n+1)^2 displaced heightfield grid, 2 n^2 triangles (config C2 uses n=2236)."""
    m = n + 1
    xs = np.arange(m, dtype=np.float64) / n
    gx, gy = np.meshgrid(xs, xs, indexing="xy")
    px = gx.ravel()
    py = gy.ravel()

    def height(x, y):
        z = np.zeros_like(x)
        for k in range(4):
            f = 4.0 * (2**k)
            pts = np.stack([x * f, y * f, np.full_like(x, 0.5 + k)], axis=1)
            z += amplitude * (0.5**k) * value_noise3(pts, seed + k)
        return z

    pz = np.empty_like(px)
    nx = np.empty_like(px)
    ny = np.empty_like(px)
    eps = 0.5 / n

    def work(s):
        e = min(px.size, s + chunk)
        x, y = px[s:e], py[s:e]
        pz[s:e] = height(x, y)
        nx[s:e] = (height(x + eps, y) - height(x - eps, y)) / (2 * eps)
        ny[s:e] = (height(x, y + eps) - height(x, y - eps)) / (2 * eps)

    # chunks are independent and numpy releases the GIL inside its loops: generate them on all host cores
    chunk = min(chunk, 1 << 18)
    starts = list(range(0, px.size, chunk))
    if len(starts) > 1:
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
            list(ex.map(work, starts))
    else:
        for s in starts:
            work(s)
    nrm = np.stack([-nx, -ny, np.ones_like(nx)], axis=1)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    verts = np.concatenate([np.stack([px, py, pz], axis=1), nrm], axis=1).astype(np.float32)

    q = np.arange(n, dtype=np.uint32)
    qx, qy = np.meshgrid(q, q, indexing="xy")
    v00 = (qy * np.uint32(m) + qx).ravel()
    v10 = v00 + np.uint32(1)
    v01 = v00 + np.uint32(m)
    v11 = v01 + np.uint32(1)
    tris = np.stack([v00, v10, v11, v00, v11, v01], axis=1).reshape(-1)
    return Mesh(verts, tris.astype(np.uint32), VERTEX_NORMALS, f"grid{n}")


_ICO_T = (1.0 + 5.0**0.5) / 2.0
_ICO_V = np.array(
    [[-1, _ICO_T, 0], [1, _ICO_T, 0], [-1, -_ICO_T, 0], [1, -_ICO_T, 0], [0, -1, _ICO_T], [0, 1, _ICO_T], [0, -1, -_ICO_T], [0, 1, -_ICO_T], [_ICO_T, 0, -1], [_ICO_T, 0, 1], [-_ICO_T, 0, -1], [-_ICO_T, 0, 1]],
    dtype=np.float64,
)
_ICO_F = np.array(
    [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]],
    dtype=np.int64,
)


def _face_lattice(f: int):
    """Integer barycentric lattice of one frequency-f face: rows i (0..f), columns j (0..f-i)."""
    ii, jj = np.meshgrid(np.arange(f + 1), np.arange(f + 1), indexing="ij")
    keep = (ii + jj) <= f
    i = ii[keep]
    j = jj[keep]
    lid = np.full((f + 1, f + 1), -1, dtype=np.int64)
    lid[i, j] = np.arange(i.size)
    # upward triangles (i,j),(i+1,j),(i,j+1) ; downward (i+1,j),(i+1,j+1),(i,j+1)
    ui, uj = np.meshgrid(np.arange(f), np.arange(f), indexing="ij")
    um = (ui + uj) < f
    ui, uj = ui[um], uj[um]
    up = np.stack([lid[ui, uj], lid[ui + 1, uj], lid[ui, uj + 1]], axis=1)
    dm = (ui + uj) < f - 1
    di, dj = ui[dm], uj[dm]
    down = np.stack([lid[di + 1, dj], lid[di + 1, dj + 1], lid[di, dj + 1]], axis=1)
    return i, j, np.concatenate([up, down], axis=0)


def icosphere(f: int, displace: bool = False, uv_atlas: bool = False, seed: int = 42) -> Mesh:
    """Frequency-f icosphere, 20 f^2 triangles.

    displace=False, uv_atlas=False : welded unit sphere, analytic normals (config C1, f=224).
    displace=True,  uv_atlas=True  : radial noise displacement, per-face UV charts whose borders duplicate
                                     positions with different UVs => attribute seams (config C3, f=2236).
    """
    bi, bj, ltris = _face_lattice(f)
    bk = f - bi - bj
    npf = bi.size
    w = np.stack([bk, bi, bj], axis=1).astype(np.float64) / f  # weights of face corners (A,B,C)
    pos_all = []
    key_all = []
    uv_all = []
    edge_ids = {}
    for fi, (a, b, c) in enumerate(_ICO_F):
        P = w[:, 0:1] * _ICO_V[a] + w[:, 1:2] * _ICO_V[b] + w[:, 2:3] * _ICO_V[c]
        pos_all.append(P)
        if uv_atlas:
            cx, cy = (fi % 5) * 0.2, (fi // 5) * 0.25
            uv_all.append(np.stack([cx + 0.01 + 0.18 * (w[:, 1] + 0.5 * w[:, 2]), cy + 0.01 + 0.23 * w[:, 2]], axis=1))
        else:
            # canonical weld keys: corner / edge / interior
            key = np.empty(npf, dtype=np.int64)
            wa, wb, wc = bk, bi, bj
            interior = (wa > 0) & (wb > 0) & (wc > 0)
            key[interior] = (1 << 40) + fi * npf + np.nonzero(interior)[0]
            for (u, v, wu, wv, wo) in ((a, b, wa, wb, wc), (b, c, wb, wc, wa), (a, c, wa, wc, wb)):
                on = (wo == 0) & (wu > 0) & (wv > 0)
                lo, hi = (u, v) if u < v else (v, u)
                eid = edge_ids.setdefault((lo, hi), len(edge_ids))
                tlo = wu if u < v else wv
                key[on] = (1 << 30) + eid * (f + 1) + tlo[on]
            for (u, wu) in ((a, wa), (b, wb), (c, wc)):
                key[wu == f] = u
            key_all.append(key)
    P = np.concatenate(pos_all, axis=0)
    tris = np.concatenate([ltris + fi * npf for fi in range(20)], axis=0)
    if not uv_atlas:
        key = np.concatenate(key_all)
        _, first, inverse = np.unique(key, return_index=True, return_inverse=True)
        P = P[first]
        tris = inverse[tris]
    P /= np.linalg.norm(P, axis=1, keepdims=True)
    if displace:
        def radius(d):
            return 1.0 + 0.02 * value_noise3(d * 8.0 + 100.0, seed) + 0.002 * value_noise3(d * 64.0 + 100.0, seed + 1)

        r = radius(P)
        # normal of r(d) d: d * r - tangential gradient of r; central differences on an orthonormal tangent frame
        eps = 1e-3
        ref = np.where(np.abs(P[:, 0:1]) < 0.9, np.array([[1.0, 0.0, 0.0]]), np.array([[0.0, 1.0, 0.0]]))
        t1 = np.cross(P, ref)
        t1 /= np.linalg.norm(t1, axis=1, keepdims=True)
        t2 = np.cross(P, t1)

        def dirn(v):
            return v / np.linalg.norm(v, axis=1, keepdims=True)

        g1 = (radius(dirn(P + eps * t1)) - radius(dirn(P - eps * t1))) / (2 * eps)
        g2 = (radius(dirn(P + eps * t2)) - radius(dirn(P - eps * t2))) / (2 * eps)
        N = P * r[:, None] - t1 * g1[:, None] - t2 * g2[:, None]
        N /= np.linalg.norm(N, axis=1, keepdims=True)
        P = P * r[:, None]
    else:
        N = P.copy()
    cols = [P, N]
    flags = VERTEX_NORMALS
    if uv_atlas:
        cols.append(np.concatenate(uv_all, axis=0))
        flags |= VERTEX_TEXCOORDS
    verts = np.concatenate(cols, axis=1).astype(np.float32)
    if uv_atlas:
        # chart borders must carry bit-identical positions/normals so that they weld by position (seams, not cracks)
        P32 = verts[:, 0:3]
        q = np.round(P.astype(np.float64) / np.linalg.norm(P, axis=1, keepdims=True) * (4.0 * f)).astype(np.int64)
        qk = (q[:, 0] + (1 << 20)) | ((q[:, 1] + (1 << 20)) << 21) | ((q[:, 2] + (1 << 20)) << 42)
        _, first, inverse = np.unique(qk, return_index=True, return_inverse=True)
        verts[:, 0:6] = verts[first[inverse], 0:6]
        del P32
    name = f"icosphere{f}" + ("_disp" if displace else "") + ("_uv" if uv_atlas else "")
    return Mesh(verts, tris.reshape(-1).astype(np.uint32), flags, name)


def torus(nu: int, nv: int, seed: int = 0, major: float = 1.0, minor: float = 0.35) -> Mesh:
    """Welded torus with 2 nu nv triangles and light noise displacement along the normal."""
    u = np.arange(nu, dtype=np.float64) / nu * 2 * np.pi
    v = np.arange(nv, dtype=np.float64) / nv * 2 * np.pi
    uu, vv = np.meshgrid(u, v, indexing="xy")
    uu, vv = uu.ravel(), vv.ravel()
    N = np.stack([np.cos(vv) * np.cos(uu), np.cos(vv) * np.sin(uu), np.sin(vv)], axis=1)
    C = np.stack([major * np.cos(uu), major * np.sin(uu), np.zeros_like(uu)], axis=1)
    P = C + minor * N
    P = P + N * (0.01 * value_noise3(P * 6.0 + 50.0, seed))[:, None]
    verts = np.concatenate([P, N], axis=1).astype(np.float32)
    iu = np.arange(nu, dtype=np.int64)
    iv = np.arange(nv, dtype=np.int64)
    a, b = np.meshgrid(iu, iv, indexing="xy")
    a, b = a.ravel(), b.ravel()
    a1, b1 = (a + 1) % nu, (b + 1) % nv
    v00, v10, v01, v11 = b * nu + a, b * nu + a1, b1 * nu + a, b1 * nu + a1
    tris = np.stack([v00, v10, v11, v00, v11, v01], axis=1).reshape(-1)
    return Mesh(verts, tris.astype(np.uint32), VERTEX_NORMALS, f"torus{nu}x{nv}")


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def scene_batch_sizes(count: int, total_tris: float, lo: float = 1e4, hi: float = 2e6, seed: int = 7):
    """Triangle budgets of a scene batch: log-uniform in [lo, hi], rescaled to total_tris (configs C4/C5)."""
    u = np.array([_splitmix64(seed * 0x100000001B3 + i) / 2.0**64 for i in range(count)])
    t = np.exp(np.log(lo) + u * (np.log(hi) - np.log(lo)))
    t *= total_tris / t.sum()
    return np.maximum(t, 1000.0)


def scene_mesh(i: int, target_tris: float) -> Mesh:
    """Mesh i of a scene batch: type = i mod 3 in {sphere, grid, torus}, per-mesh noise seed = i."""
    kind = i % 3
    if kind == 0:
        f = max(2, int(round((target_tris / 20.0) ** 0.5)))
        return icosphere(f, displace=True, uv_atlas=False, seed=i)
    if kind == 1:
        n = max(2, int(round((target_tris / 2.0) ** 0.5)))
        return grid(n, seed=i)
    nu = max(3, int(round((target_tris / 2.0 * 2.0) ** 0.5)))
    nv = max(3, int(round(target_tris / 2.0 / nu)))
    return torus(nu, nv, seed=i)


# ---------------------------------------------------------------------------------------------------------------------
# Config C3 at full size (f = 2236: 99 993 920 triangles, 50 M wedges). The numpy generator above needs minutes and tens of
# GB of float64 temporaries at that size; this is the same construction written with torch ops so that it can run on the
# GPU that is about to build the mesh. Deterministic; on CPU it reproduces icosphere(f, True, True) (tests/test_meshgen.py).


def _t_hash_u32(x):
    m = 0xFFFFFFFF
    x = x & m
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & m
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & m
    x = x ^ (x >> 16)
    return x


def _t_lattice(ix, iy, iz, seed):
    m = 0xFFFFFFFF
    h = _t_hash_u32((((ix & m) * 0x9E3779B1) & m) ^ _t_hash_u32((((iy & m) * 0x85EBCA77) & m) ^ _t_hash_u32((((iz & m) * 0xC2B2AE3D) & m) ^ (seed & m))))
    return (h >> 8).double() * (1.0 / 16777216.0) * 2.0 - 1.0


def _t_value_noise3(p, seed):
    import torch

    pf = torch.floor(p)
    t = p - pf
    t = t * t * (3.0 - 2.0 * t)
    i = pf.long()
    ix, iy, iz = i[:, 0], i[:, 1], i[:, 2]
    out = torch.zeros(p.shape[0], dtype=torch.float64, device=p.device)
    for dx in (0, 1):
        wx = t[:, 0] if dx else 1.0 - t[:, 0]
        for dy in (0, 1):
            wy = t[:, 1] if dy else 1.0 - t[:, 1]
            for dz in (0, 1):
                wz = t[:, 2] if dz else 1.0 - t[:, 2]
                out += wx * wy * wz * _t_lattice(ix + dx, iy + dy, iz + dz, seed)
    return out


def icosphere_seams_torch(f: int, seed: int = 42, device=None, chunk: int = 1 << 23):
    """icosphere(f, displace=True, uv_atlas=True) built with torch on `device` (default: cuda if available).

    Returns (Mesh, tangents[V,4] f32). The tangent stream is analytic — the chart's dP/du projected onto the tangent plane
    of the vertex normal, w = +1 — and differs across chart borders like the UVs do; it stands in for the MikkTSpace
    stream the reference generates internally (clodb200_geometry::tangents, include/clodb200.h)."""
    import torch

    dev = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    f64 = torch.float64
    ar = torch.arange(f + 1, device=dev)
    ii = ar.view(-1, 1).expand(f + 1, f + 1)
    jj = ar.view(1, -1).expand(f + 1, f + 1)
    keep = (ii + jj) <= f
    bi, bj = ii[keep], jj[keep]
    npf = bi.numel()
    lid = torch.full((f + 1, f + 1), -1, dtype=torch.int64, device=dev)
    lid[bi, bj] = torch.arange(npf, device=dev)
    af = torch.arange(f, device=dev)
    ui = af.view(-1, 1).expand(f, f)
    uj = af.view(1, -1).expand(f, f)
    um = (ui + uj) < f
    ui_, uj_ = ui[um], uj[um]
    up = torch.stack([lid[ui_, uj_], lid[ui_ + 1, uj_], lid[ui_, uj_ + 1]], dim=1)
    dm = (ui + uj) < f - 1
    di, dj = ui[dm], uj[dm]
    down = torch.stack([lid[di + 1, dj], lid[di + 1, dj + 1], lid[di, dj + 1]], dim=1)
    ltris = torch.cat([up, down], dim=0)
    del lid, up, down, ui_, uj_, di, dj
    bk = f - bi - bj
    w = torch.stack([bk, bi, bj], dim=1).to(f64) / f
    icov = torch.tensor(_ICO_V, dtype=f64, device=dev)

    V = npf * 20
    verts = torch.empty((V, 8), dtype=torch.float32, device=dev)
    tang = torch.empty((V, 4), dtype=torch.float32, device=dev)
    qk = torch.empty(V, dtype=torch.int64, device=dev)

    def radius(d):
        return 1.0 + 0.02 * _t_value_noise3(d * 8.0 + 100.0, seed) + 0.002 * _t_value_noise3(d * 64.0 + 100.0, seed + 1)

    def dirn(v):
        return v / torch.linalg.norm(v, dim=1, keepdim=True)

    ex = torch.tensor([[1.0, 0.0, 0.0]], dtype=f64, device=dev)
    ey = torch.tensor([[0.0, 1.0, 0.0]], dtype=f64, device=dev)
    eps = 1e-3
    for fi, (a, b, c) in enumerate(_ICO_F):
        for s in range(0, npf, chunk):
            e = min(npf, s + chunk)
            ws = w[s:e]
            P = ws[:, 0:1] * icov[a] + ws[:, 1:2] * icov[b] + ws[:, 2:3] * icov[c]
            P = P / torch.linalg.norm(P, dim=1, keepdim=True)
            r = radius(P)
            ref = torch.where(P[:, 0:1].abs() < 0.9, ex, ey)
            t1 = torch.linalg.cross(P, ref.expand_as(P))
            t1 = t1 / torch.linalg.norm(t1, dim=1, keepdim=True)
            t2 = torch.linalg.cross(P, t1)
            g1 = (radius(dirn(P + eps * t1)) - radius(dirn(P - eps * t1))) / (2 * eps)
            g2 = (radius(dirn(P + eps * t2)) - radius(dirn(P - eps * t2))) / (2 * eps)
            N = P * r[:, None] - t1 * g1[:, None] - t2 * g2[:, None]
            N = N / torch.linalg.norm(N, dim=1, keepdim=True)
            Pd = P * r[:, None]
            o = fi * npf
            verts[o + s : o + e, 0:3] = Pd.float()
            verts[o + s : o + e, 3:6] = N.float()
            cx, cy = (fi % 5) * 0.2, (fi // 5) * 0.25
            verts[o + s : o + e, 6] = (cx + 0.01 + 0.18 * (ws[:, 1] + 0.5 * ws[:, 2])).float()
            verts[o + s : o + e, 7] = (cy + 0.01 + 0.23 * ws[:, 2]).float()
            q = torch.round(Pd / torch.linalg.norm(Pd, dim=1, keepdim=True) * (4.0 * f)).long()
            qk[o + s : o + e] = (q[:, 0] + (1 << 20)) | ((q[:, 1] + (1 << 20)) << 21) | ((q[:, 2] + (1 << 20)) << 42)
            # chart u direction (towards corner b), made tangent to the shaded surface
            du = (icov[b] - icov[a]).view(1, 3)
            T = du - N * (N * du).sum(dim=1, keepdim=True)
            T = T / torch.linalg.norm(T, dim=1, keepdim=True)
            tang[o + s : o + e, 0:3] = T.float()
            tang[o + s : o + e, 3] = 1.0
            del P, r, t1, t2, g1, g2, N, Pd, q, T
    # chart borders must carry bit-identical positions/normals so that they weld by position (seams, not cracks)
    _, inverse = torch.unique(qk, return_inverse=True)
    first = torch.full((int(inverse.max()) + 1,), V, dtype=torch.int64, device=dev)
    first.scatter_reduce_(0, inverse, torch.arange(V, device=dev), reduce="amin")
    src = first[inverse]
    del qk, inverse, first
    verts[:, 0:6] = verts[src, 0:6]
    del src
    tris = torch.cat([ltris + fi * npf for fi in range(20)], dim=0).reshape(-1).to(torch.int32)
    mesh = Mesh(verts.cpu().numpy(), tris.cpu().numpy().view(np.uint32), VERTEX_NORMALS | VERTEX_TEXCOORDS, f"icosphere{f}_disp_uv")
    return mesh, tang.cpu().numpy()


# ---------------------------------------------------------------------------------------------------------------------
# Scene batches at full size (configs C4 / C5: 512 / 4 096 meshes, ~1 B / ~8 B triangles): the same three generators
# written with torch ops so that a rank can generate its shard on its GPU in seconds (the numpy versions evaluate the value noise
# on one host core). Same formulas and seeds as grid() / torus() / icosphere(displace=True); float operation order on the GPU is
# not the CPU's, so these are the same surfaces, not bit-identical vertex streams.


def _t_device(device):
    import torch

    return torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")


def _t_grid_indices(nx_quads: int, ny_quads: int, row: int, dev, wrap_x: bool = False, wrap_y: bool = False):
    import torch

    a = torch.arange(nx_quads, device=dev).view(1, -1).expand(ny_quads, nx_quads).reshape(-1)
    b = torch.arange(ny_quads, device=dev).view(-1, 1).expand(ny_quads, nx_quads).reshape(-1)
    a1 = (a + 1) % nx_quads if wrap_x else a + 1
    b1 = (b + 1) % ny_quads if wrap_y else b + 1
    v00, v10, v01, v11 = b * row + a, b * row + a1, b1 * row + a, b1 * row + a1
    return torch.stack([v00, v10, v11, v00, v11, v01], dim=1).reshape(-1).to(torch.int32)


def grid_torch(n: int, seed: int = 1234, amplitude: float = 0.1, device=None) -> Mesh:
    import torch

    dev = _t_device(device)
    m = n + 1
    xs = torch.arange(m, dtype=torch.float64, device=dev) / n
    px = xs.view(1, -1).expand(m, m).reshape(-1)
    py = xs.view(-1, 1).expand(m, m).reshape(-1)

    def height(x, y):
        z = torch.zeros_like(x)
        for k in range(4):
            f = 4.0 * (2**k)
            pts = torch.stack([x * f, y * f, torch.full_like(x, 0.5 + k)], dim=1)
            z += amplitude * (0.5**k) * _t_value_noise3(pts, seed + k)
        return z

    eps = 0.5 / n
    pz = height(px, py)
    nx = (height(px + eps, py) - height(px - eps, py)) / (2 * eps)
    ny = (height(px, py + eps) - height(px, py - eps)) / (2 * eps)
    nrm = torch.stack([-nx, -ny, torch.ones_like(nx)], dim=1)
    nrm = nrm / torch.linalg.norm(nrm, dim=1, keepdim=True)
    verts = torch.cat([torch.stack([px, py, pz], dim=1), nrm], dim=1).float()
    tris = _t_grid_indices(n, n, m, dev)
    return Mesh(verts.cpu().numpy(), tris.cpu().numpy().view(np.uint32), VERTEX_NORMALS, f"grid{n}")


def torus_torch(nu: int, nv: int, seed: int = 0, major: float = 1.0, minor: float = 0.35, device=None) -> Mesh:
    import math

    import torch

    dev = _t_device(device)
    u = torch.arange(nu, dtype=torch.float64, device=dev) / nu * 2 * math.pi
    v = torch.arange(nv, dtype=torch.float64, device=dev) / nv * 2 * math.pi
    uu = u.view(1, -1).expand(nv, nu).reshape(-1)
    vv = v.view(-1, 1).expand(nv, nu).reshape(-1)
    N = torch.stack([torch.cos(vv) * torch.cos(uu), torch.cos(vv) * torch.sin(uu), torch.sin(vv)], dim=1)
    Cc = torch.stack([major * torch.cos(uu), major * torch.sin(uu), torch.zeros_like(uu)], dim=1)
    P = Cc + minor * N
    P = P + N * (0.01 * _t_value_noise3(P * 6.0 + 50.0, seed)).unsqueeze(1)
    verts = torch.cat([P, N], dim=1).float()
    tris = _t_grid_indices(nu, nv, nu, dev, wrap_x=True, wrap_y=True)
    return Mesh(verts.cpu().numpy(), tris.cpu().numpy().view(np.uint32), VERTEX_NORMALS, f"torus{nu}x{nv}")


def icosphere_torch(f: int, seed: int = 42, device=None) -> Mesh:
    """Welded, displaced icosphere with gradient normals (icosphere(f, displace=True, uv_atlas=False)): the seamed construction with
    the chart-border duplicates merged by position."""
    import torch

    dev = _t_device(device)
    mesh, _ = icosphere_seams_torch(f, seed=seed, device=dev)
    v = torch.from_numpy(mesh.vertices).to(dev)
    idx = torch.from_numpy(mesh.indices.view(np.int32).astype(np.int64)).to(dev)
    key = v[:, 0:3].contiguous().view(torch.int32).long()  # chart borders carry bit-identical positions
    packed = (key[:, 0] & 0xFFFFFFFF) * 0x9E3779B1 ^ (key[:, 1] & 0xFFFFFFFF) * 0x85EBCA77 ^ (key[:, 2] & 0xFFFFFFFF) * 0xC2B2AE3D
    # exact weld: unique rows of the three position words
    _, inverse = torch.unique(key, dim=0, return_inverse=True)
    del packed
    count = int(inverse.max()) + 1
    first = torch.full((count,), v.shape[0], dtype=torch.int64, device=dev)
    first.scatter_reduce_(0, inverse, torch.arange(v.shape[0], device=dev), reduce="amin")
    order = torch.argsort(first)  # keep first-occurrence vertex order
    rank = torch.empty_like(order)
    rank[order] = torch.arange(count, device=dev)
    verts = v[first[order], 0:6].contiguous()
    tris = rank[inverse[idx]].to(torch.int32)
    return Mesh(verts.cpu().numpy(), tris.cpu().numpy().view(np.uint32), VERTEX_NORMALS, f"icosphere{f}_disp")


def scene_mesh_torch(i: int, target_tris: float, device=None) -> Mesh:
    """scene_mesh(i, target_tris) generated with torch on `device` (same type rule, sizes and seeds)."""
    kind = i % 3
    if kind == 0:
        f = max(2, int(round((target_tris / 20.0) ** 0.5)))
        return icosphere_torch(f, seed=i, device=device)
    if kind == 1:
        n = max(2, int(round((target_tris / 2.0) ** 0.5)))
        return grid_torch(n, seed=i, device=device)
    nu = max(3, int(round((target_tris / 2.0 * 2.0) ** 0.5)))
    nv = max(3, int(round(target_tris / 2.0 / nu)))
    return torus_torch(nu, nv, seed=i, device=device)
