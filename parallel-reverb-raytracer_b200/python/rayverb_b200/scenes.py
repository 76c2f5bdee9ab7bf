"""Scene and ray generators for the configurations named in BASELINE.json / SURVEY.md §8(d).

Everything here produces the reference's input arrays (Triangle[], cl_float3 vertices[],
Surface[]; rayverb/clstructs.h:17-32) -- the geometry itself is synthetic, deterministic and
independent of the reference tree.
"""
import json
import math

import numpy as np

from .types import SURFACE, TRIANGLE

# The loader's hard-coded default surface (rayverb.cpp:336-339)
DEFAULT_SPECULAR = (0.92, 0.92, 0.93, 0.93, 0.94, 0.95, 0.95, 0.95)
DEFAULT_DIFFUSE = (0.50, 0.90, 0.95, 0.95, 0.95, 0.95, 0.95, 0.95)


def make_surfaces(pairs):
    s = np.zeros(len(pairs), dtype=SURFACE)
    for i, (spec, diff) in enumerate(pairs):
        s["specular"][i] = spec
        s["diffuse"][i] = diff
    return s


def default_surfaces():
    return make_surfaces([(DEFAULT_SPECULAR, DEFAULT_DIFFUSE)])


def _pack(vertices, faces, surface_ids):
    v = np.zeros((len(vertices), 4), dtype=np.float32)
    v[:, :3] = np.asarray(vertices, dtype=np.float32).reshape(-1, 3)
    t = np.zeros(len(faces), dtype=TRIANGLE)
    f = np.asarray(faces, dtype=np.uint64).reshape(-1, 3)
    t["v0"], t["v1"], t["v2"] = f[:, 0], f[:, 1], f[:, 2]
    t["surface"] = np.asarray(surface_ids, dtype=np.uint64)
    return t, v


def shoebox(lx=10.0, ly=4.0, lz=6.0):
    """C2: axis-aligned box [0,lx]x[0,ly]x[0,lz], 2 triangles per face, default surface."""
    c = [(x, y, z) for x in (0.0, lx) for y in (0.0, ly) for z in (0.0, lz)]  # index = 4*ix+2*iy+iz
    quads = [(0, 1, 3, 2), (4, 6, 7, 5),      # x = 0, x = lx
             (0, 4, 5, 1), (2, 3, 7, 6),      # y = 0, y = ly
             (0, 2, 6, 4), (1, 5, 7, 3)]      # z = 0, z = lz
    faces = []
    for a, b, cc, d in quads:
        faces += [(a, b, cc), (a, cc, d)]
    tri, vtx = _pack(c, faces, [0] * len(faces))
    return tri, vtx, default_surfaces()


def diamond_room():
    """The 27 m diamond-plan room of the reference's RaytracerTest fixture
    (tests/raytrace_tests.cpp:6-25 loads demo/assets/test_models/large_square.obj):
    six quads, un-joined vertices per face, fan-triangulated, default surface
    (its material name is not in mat.json -> surface 0, rayverb.cpp:369-372)."""
    e1, e2, e3 = -3.30655e-15, 1.65327e-15, -4.95982e-15
    v = [(-27, 0, e1), (e2, 0, -27), (27, 0, -0.0), (e3, 0, 27),
         (-27, 27, e1), (e2, 27, -27), (e3, 27, 27), (27, 27, -0.0)]
    quads = [(1, 2, 3, 4), (2, 1, 5, 6), (1, 4, 7, 5), (4, 3, 8, 7), (3, 2, 6, 8), (6, 5, 7, 8)]
    verts, faces = [], []
    for q in quads:
        base = len(verts)
        verts += [v[i - 1] for i in q]
        faces += [(base, base + 1, base + 2), (base, base + 2, base + 3)]
    tri, vtx = _pack(verts, faces, [0] * len(faces))
    return tri, vtx, default_surfaces()


def _grid_face(origin, du, dv, nu, nv, normal, amp, pu, pv, warp):
    """(nu x nv) quad grid spanning origin + a*du + b*dv, displaced along `normal` by
    amp*sin(pi*pu*a)*sin(pi*pv*b) (exactly zero on the rim so neighbouring faces stay joined)."""
    origin, du, dv, normal = (np.asarray(x, dtype=np.float64) for x in (origin, du, dv, normal))
    ii, jj = np.meshgrid(np.arange(nu + 1), np.arange(nv + 1), indexing="ij")
    a = ii / nu
    b = jj / nv
    sa = np.sin(math.pi * pu * a)
    sb = np.sin(math.pi * pv * b)
    sa[(ii == 0) | (ii == nu)] = 0.0
    sb[(jj == 0) | (jj == nv)] = 0.0
    p = origin + a[..., None] * du + b[..., None] * dv + (amp * sa * sb)[..., None] * normal
    p = warp(p.reshape(-1, 3))
    idx = (ii * (nv + 1) + jj)
    q00 = idx[:-1, :-1].ravel(); q10 = idx[1:, :-1].ravel(); q11 = idx[1:, 1:].ravel(); q01 = idx[:-1, 1:].ravel()
    faces = np.empty((2 * nu * nv, 3), dtype=np.int64)
    faces[0::2] = np.stack([q00, q10, q11], 1)
    faces[1::2] = np.stack([q00, q11, q01], 1)
    return p, faces


def displaced_box(lx, ly, lz, target_triangles, amp=0.05, fan=0.0, nsurfaces=1, vault=0.0):
    """Closed room: a box whose six walls are regular grids displaced by a small sinusoid.
    fan > 0 widens x with z (fan-shaped plan); vault > 0 bulges the ceiling into a barrel vault."""
    area = 2 * (lx * lz + lx * ly + ly * lz)
    h = math.sqrt(2.0 * area / target_triangles)
    nx, ny, nz = (max(1, round(l / h)) for l in (lx, ly, lz))

    def warp(p):
        q = p.copy()
        if vault > 0.0:  # raise y near the top following a circular arc over x
            top = np.clip(p[:, 1] / ly, 0.0, 1.0) ** 4
            q[:, 1] = p[:, 1] + vault * top * np.sin(math.pi * np.clip(p[:, 0] / lx, 0.0, 1.0))
        if fan > 0.0:
            q[:, 0] = (p[:, 0] - 0.5 * lx) * (1.0 + fan * p[:, 2] / lz) + 0.5 * lx
        return q

    specs = [  # origin, du, dv, nu, nv, inward normal
        ((0, 0, 0), (lx, 0, 0), (0, 0, lz), nx, nz, (0, 1, 0)),      # floor
        ((0, ly, 0), (lx, 0, 0), (0, 0, lz), nx, nz, (0, -1, 0)),    # ceiling
        ((0, 0, 0), (0, ly, 0), (0, 0, lz), ny, nz, (1, 0, 0)),      # x = 0
        ((lx, 0, 0), (0, ly, 0), (0, 0, lz), ny, nz, (-1, 0, 0)),    # x = lx
        ((0, 0, 0), (lx, 0, 0), (0, ly, 0), nx, ny, (0, 0, 1)),      # z = 0
        ((0, 0, lz), (lx, 0, 0), (0, ly, 0), nx, ny, (0, 0, -1)),    # z = lz
    ]
    verts, faces, sids = [], [], []
    base = 0
    for k, (o, du, dv, nu, nv, nrm) in enumerate(specs):
        p, f = _grid_face(o, du, dv, nu, nv, nrm, amp, 3 + k, 2 + k, warp)
        verts.append(p); faces.append(f + base); base += len(p)
        sids.append(np.full(len(f), k % nsurfaces, dtype=np.int64))
    tri, vtx = _pack(np.concatenate(verts), np.concatenate(faces), np.concatenate(sids))
    return tri, vtx


def _box_mesh(lo, hi, n):
    """Closed axis-aligned box [lo, hi], every face an n x n quad grid (12 n^2 triangles)."""
    lo = np.asarray(lo, dtype=np.float64); hi = np.asarray(hi, dtype=np.float64)
    d = hi - lo
    ex, ey, ez = np.array([d[0], 0, 0]), np.array([0, d[1], 0]), np.array([0, 0, d[2]])
    faces_spec = [(lo, ey, ez), (lo + ex, ey, ez), (lo, ex, ez), (lo + ey, ex, ez), (lo, ex, ey), (lo + ez, ex, ey)]
    verts, faces, base = [], [], 0
    ident = lambda p: p  # noqa: E731
    for o, du, dv in faces_spec:
        p, f = _grid_face(o, du, dv, n, n, (0, 0, 0), 0.0, 1, 1, ident)
        verts.append(p); faces.append(f + base); base += len(p)
    return np.concatenate(verts), np.concatenate(faces)


def pillar_room(target_triangles=4000, npillars=6, pillar_grid=6):
    """A displaced box with free-standing pillars between typical source / microphone positions:
    unlike the closed halls, many hit points here do NOT see the microphone (occluded visibility
    queries, kernel.cpp:463-471) and rays reach the pillars from all sides."""
    room_tris = max(200, target_triangles - npillars * 12 * pillar_grid * pillar_grid)
    tri, vtx = displaced_box(16.0, 6.0, 12.0, room_tris, amp=0.06, fan=0.15, nsurfaces=4)
    verts = [vtx[:, :3].astype(np.float64)]
    faces = [np.stack([tri["v0"], tri["v1"], tri["v2"]], 1).astype(np.int64)]
    sids = [tri["surface"].astype(np.int64)]
    base = len(vtx)
    for k in range(npillars):
        cx = 3.0 + 2.0 * k
        cz = 4.0 + (1.7 * k) % 4.5
        w = 0.35 + 0.05 * (k % 3)
        p, f = _box_mesh((cx - w, 0.02, cz - w), (cx + w, 5.2 - 0.3 * (k % 2), cz + w), pillar_grid)
        verts.append(p); faces.append(f + base); base += len(p)
        sids.append(np.full(len(f), (k + 1) % 4, dtype=np.int64))
    t, v = _pack(np.concatenate(verts), np.concatenate(faces), np.concatenate(sids))
    return t, v, hall_surfaces()


def hall_surfaces():
    """Four surfaces in the value ranges of the shipped material files (specular 0.12-0.99,
    diffuse 0.6-0.99)."""
    return make_surfaces([
        (DEFAULT_SPECULAR, DEFAULT_DIFFUSE),
        ((0.90, 0.94, 0.95, 0.95, 0.95, 0.96, 0.96, 0.96), (0.70, 0.90, 0.90, 0.90, 0.90, 0.80, 0.80, 0.70)),
        ((0.80, 0.88, 0.90, 0.91, 0.92, 0.93, 0.93, 0.93), (0.60, 0.80, 0.85, 0.90, 0.90, 0.90, 0.85, 0.80)),
        ((0.50, 0.45, 0.40, 0.35, 0.30, 0.25, 0.20, 0.12), (0.99, 0.95, 0.90, 0.85, 0.80, 0.75, 0.70, 0.60)),
    ])


def concert_hall(target_triangles=100_000):
    """C3: fan-shaped hall ~40 x 15 x 25 m, ~100k triangles, 4 surfaces cycled by wall."""
    tri, vtx = displaced_box(40.0, 15.0, 25.0, target_triangles, amp=0.05, fan=0.3, nsurfaces=4)
    return tri, vtx, hall_surfaces()


def cathedral(target_triangles=1_000_000):
    """C4: long nave 30 x 25 x 90 m with a barrel-vaulted ceiling, ~1M triangles (edge ~0.1 m)."""
    tri, vtx = displaced_box(30.0, 25.0, 90.0, target_triangles, amp=0.03, fan=0.0, nsurfaces=4, vault=8.0)
    return tri, vtx, hall_surfaces()


def sphere_directions(n, seed=2015):
    """n unit vectors: z ~ U[-1,1), theta ~ U[-pi,pi) drawn (z first) from MT19937(seed), mapped
    through spherePoint (helpers.cpp:63-67) in float32."""
    rng = np.random.Generator(np.random.MT19937(seed))
    zt = rng.random((n, 2), dtype=np.float32)
    z = (zt[:, 0] * np.float32(2.0) - np.float32(1.0)).astype(np.float32)
    theta = ((zt[:, 1] * np.float32(2.0) - np.float32(1.0)) * np.float32(math.pi)).astype(np.float32)
    ztemp = np.sqrt(np.float32(1.0) - z * z, dtype=np.float32)
    d = np.zeros((n, 4), dtype=np.float32)
    d[:, 0] = ztemp * np.cos(theta, dtype=np.float32)
    d[:, 1] = ztemp * np.sin(theta, dtype=np.float32)
    d[:, 2] = z
    return d


# ---------------------------------------------------------------------------------------------
# OBJ + materials loader with the observable behaviour of the reference's SceneData
# (rayverb.cpp:296-461): default surface at index 0, materials sorted by name at 1.., one mesh
# per `usemtl` run, vertices NOT joined across faces, polygons triangulated.
# ---------------------------------------------------------------------------------------------
def _ear_clip(poly3d):
    """Triangulate a planar simple polygon (list of 3-D points); returns index triples."""
    n = len(poly3d)
    if n == 3:
        return [(0, 1, 2)]
    p = np.asarray(poly3d, dtype=np.float64)
    nrm = np.zeros(3)
    for i in range(n):  # Newell normal
        a, b = p[i], p[(i + 1) % n]
        nrm += np.array([(a[1] - b[1]) * (a[2] + b[2]), (a[2] - b[2]) * (a[0] + b[0]), (a[0] - b[0]) * (a[1] + b[1])])
    ax = int(np.argmax(np.abs(nrm)))
    q = p[:, [(ax + 1) % 3, (ax + 2) % 3]]     # cyclic order keeps the polygon's orientation
    if nrm[ax] < 0:
        q = q[:, ::-1]

    def cross2(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    idx = list(range(n))
    out = []
    guard = 0
    while len(idx) > 3 and guard < 10 * n:
        guard += 1
        m = len(idx)
        clipped = False
        for kk in range(m):
            k = (kk + 1) % m     # start at corner 1: a convex polygon comes out as a fan around vertex 0
            i0, i1, i2 = idx[(k - 1) % m], idx[k], idx[(k + 1) % m]
            if cross2(q[i0], q[i1], q[i2]) <= 1e-12:
                continue
            inside = False
            for j in idx:
                if j in (i0, i1, i2):
                    continue
                if cross2(q[i0], q[i1], q[j]) >= 0 and cross2(q[i1], q[i2], q[j]) >= 0 and cross2(q[i2], q[i0], q[j]) >= 0:
                    inside = True
                    break
            if inside:
                continue
            out.append((i0, i1, i2))
            del idx[k]
            clipped = True
            break
        if not clipped:  # degenerate: fall back to a fan over what is left
            break
    for k in range(1, len(idx) - 1):
        out.append((idx[0], idx[k], idx[k + 1]))
    return out


def load_materials(path):
    with open(path) as f:
        doc = json.load(f)
    if not isinstance(doc, dict):
        raise RuntimeError("Materials must be stored in a JSON object")
    names = sorted(doc)  # std::map order (rayverb.cpp:348-354)
    pairs = [(DEFAULT_SPECULAR, DEFAULT_DIFFUSE)] + [(doc[n]["specular"], doc[n]["diffuse"]) for n in names]
    return make_surfaces(pairs), {n: i + 1 for i, n in enumerate(names)}


def load_obj(obj_path, materials_path):
    surfaces, mat_index = load_materials(materials_path)
    pos = []
    verts, faces, sids = [], [], []
    current = 0
    with open(obj_path) as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                pos.append(tuple(float(x) for x in tok[1:4]))
            elif tok[0] == "usemtl":
                current = mat_index.get(tok[1], 0)
            elif tok[0] == "f":
                ids = []
                for t in tok[1:]:
                    i = int(t.split("/")[0])
                    ids.append(i - 1 if i > 0 else len(pos) + i)
                pts = [pos[i] for i in ids]
                base = len(verts)
                verts += pts
                for (a, b, c) in _ear_clip(pts):
                    faces.append((base + a, base + b, base + c))
                    sids.append(current)
    tri, vtx = _pack(verts, faces, sids)
    return tri, vtx, surfaces


def synthetic_hrtf_table():
    """Smooth synthetic [2, 360, 180, 8] HRTF table (the reference's rayverb/hrtf.cpp, generated from
    the IRCAM Listen recordings, is not available): positive, ear-asymmetric, band-dependent.  Same
    formula as the C++ host's built-in table (host/rayverb.cpp: fillSynthetic)."""
    a = np.deg2rad(np.arange(360, dtype=np.float64))[None, :, None, None]
    e = np.deg2rad(np.arange(180, dtype=np.float64))[None, None, :, None]
    b = np.arange(8, dtype=np.float64)[None, None, None, :]
    ear = np.array([-1.0, 1.0])[:, None, None, None]
    t = 0.55 + 0.35 * ear * np.sin(a) * np.sin(e) * (0.3 + 0.1 * b) + 0.08 * np.cos(2 * a + 0.3 * b) * np.sin(e)
    return np.ascontiguousarray(t, dtype=np.float32)
