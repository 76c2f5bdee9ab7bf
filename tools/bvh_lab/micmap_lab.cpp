// CPU lab for the microphone visibility map (development aid, not part of the product).
// Builds the cube map of triangle lists around a microphone exactly the way csrc/micmap.cu does
// and checks, with the canonical float arithmetic of the hot path (explicit fmaf, no contraction),
// that every triangle the exact Moller-Trumbore test accepts on a segment X -> mic is among the
// candidates of X's cell.  Also reports how far outside its triangle an accepted hit can lie.
//   g++ -O2 -fopenmp -ffp-contract=off -std=c++17 -o /tmp/mlab tools/bvh_lab/micmap_lab.cpp
//   /tmp/mlab /tmp/lab_hall [micx micy micz] [delta_rel] [cells_per_tri]
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

struct V3 { float x, y, z; };
static inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline float dotf(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
static inline V3 crossf(V3 a, V3 b) { return {fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))}; }
static inline float lengthf(V3 a) { return sqrtf(dotf(a, a)); }
static inline V3 normalizef(V3 a) { float l = lengthf(a); if (l == 0) return {0, 0, 0}; float r = 1.0f / l; return {a.x * r, a.y * r, a.z * r}; }
struct Tri { V3 a, b, c; };
static std::vector<Tri> tris;

// rvb_math.cuh: mt_intersect_pred
static float mt(const Tri &t, V3 pos, V3 dir, bool &ok)
{
    V3 e0 = sub(t.b, t.a), e1 = sub(t.c, t.a);
    V3 p = crossf(dir, e1);
    float det = dotf(e0, p);
    float inv = 1.0f / det;
    V3 tv = sub(pos, t.a);
    float u = inv * dotf(tv, p);
    V3 q = crossf(tv, e0);
    float v = inv * dotf(dir, q);
    float tt = inv * dotf(e1, q);
    bool rej = (-1e-4f < det && det < 1e-4f) | (u < 0.0f) | (v < 0.0f) | (1.0f < v + u);
    ok = !rej;
    return tt;
}

// ---- the map, as the CUDA build computes it -----------------------------------------------------------
struct Rect { int i0, i1, j0, j1; int np; float pu[4], pv[4], margin; };
static bool g_exact = true;
// is cell (i, j) certainly outside the projected polygon grown by r.margin?
static bool cell_outside(const Rect &r, int i, int j, int N)
{
    if (!g_exact || r.np < 3) return false;
    const float h = 2.0f / N;
    const float cu0 = -1.0f + i * h, cu1 = cu0 + h, cv0 = -1.0f + j * h, cv1 = cv0 + h;
    // orientation
    float area = 0;
    for (int k = 0; k < r.np; ++k) { const int k1 = (k + 1) % r.np; area += r.pu[k] * r.pv[k1] - r.pu[k1] * r.pv[k]; }
    if (fabsf(area) < 1e-12f) return false;
    const float sgn = area > 0 ? 1.0f : -1.0f;
    for (int k = 0; k < r.np; ++k) {
        const int k1 = (k + 1) % r.np;
        // outward normal of edge k -> k1 (for a counter-clockwise polygon: (dv, -du))
        const float du = r.pu[k1] - r.pu[k], dv = r.pv[k1] - r.pv[k];
        const float nu = sgn * dv, nv = -sgn * du;
        const float len = sqrtf(nu * nu + nv * nv);
        if (len == 0) continue;
        const float cu = nu > 0 ? cu0 : cu1, cv = nv > 0 ? cv0 : cv1;   // the cell corner deepest inside
        if (nu * (cu - r.pu[k]) + nv * (cv - r.pv[k]) > (r.margin + 1e-5f) * len * 1.0001f + 1e-7f) return true;
    }
    return false;
}
static bool face_rect(const V3 p[3], int face, float wc, float delta, int N, Rect &r)
{
    r.np = 0; r.margin = 0;
    const int a = face >> 1;
    const float s = (face & 1) ? -1.0f : 1.0f;
    float w[3], ub[3], vc[3];
    for (int k = 0; k < 3; ++k) {
        const float c[3] = {p[k].x, p[k].y, p[k].z};
        w[k] = s * c[a]; ub[k] = c[(a + 1) % 3]; vc[k] = c[(a + 2) % 3];
    }
    if (w[0] < wc && w[1] < wc && w[2] < wc) return false;
    float ulo = 3e38f, uhi = -3e38f, vlo = 3e38f, vhi = -3e38f;
    auto add = [&](float W, float U, float Vv) {
        const float wn = W - delta, wf = W + delta;   // wn > 0 because wc > delta
        const float u0 = U - delta, u1 = U + delta, v0 = Vv - delta, v1 = Vv + delta;
        const float a0 = std::min(u0 / wn, u0 / wf), a1 = std::max(u1 / wn, u1 / wf), b0 = std::min(v0 / wn, v0 / wf), b1 = std::max(v1 / wn, v1 / wf);
        ulo = std::min(ulo, a0); uhi = std::max(uhi, a1); vlo = std::min(vlo, b0); vhi = std::max(vhi, b1);
        const float pu = U / W, pv = Vv / W;
        if (r.np < 4) { r.pu[r.np] = pu; r.pv[r.np] = pv; ++r.np; }
        // the grown vertex projects inside [a0, a1] x [b0, b1]: its distance from (pu, pv) is at most the box's half diagonal
        const float mu = std::max(a1 - pu, pu - a0), mv = std::max(b1 - pv, pv - b0);
        r.margin = std::max(r.margin, sqrtf(mu * mu + mv * mv));
    };
    for (int k = 0; k < 3; ++k) {
        const int k1 = (k + 1) % 3;
        if (w[k] >= wc) add(w[k], ub[k], vc[k]);
        if ((w[k] < wc) != (w[k1] < wc)) {
            const float t = (wc - w[k]) / (w[k1] - w[k]);
            add(wc, ub[k] + t * (ub[k1] - ub[k]), vc[k] + t * (vc[k1] - vc[k]));
        }
    }
    const float m = 1e-5f;
    ulo -= m; uhi += m; vlo -= m; vhi += m;
    if (uhi < -1.0f || ulo > 1.0f || vhi < -1.0f || vlo > 1.0f) return false;
    const float h = 0.5f * N;
    auto cell = [&](float x) { float f = floorf((x + 1.0f) * h); return (int)std::min((float)(N - 1), std::max(0.0f, f)); };
    r.i0 = cell(std::max(ulo, -1.0f)); r.i1 = cell(std::min(uhi, 1.0f));
    r.j0 = cell(std::max(vlo, -1.0f)); r.j1 = cell(std::min(vhi, 1.0f));
    return true;
}

static int query_cell(V3 w, int N)
{
    const float ax = fabsf(w.x), ay = fabsf(w.y), az = fabsf(w.z);
    int a = 0; float m = ax;
    if (ay > m) { a = 1; m = ay; }
    if (az > m) { a = 2; m = az; }
    const float c[3] = {w.x, w.y, w.z};
    const int face = 2 * a + (c[a] < 0.0f ? 1 : 0);
    const float u = c[(a + 1) % 3] / m, v = c[(a + 2) % 3] / m;
    const float h = 0.5f * N;
    auto cell = [&](float x) { float f = floorf((x + 1.0f) * h); return (int)std::min((float)(N - 1), std::max(0.0f, f)); };
    return (face * N + cell(v)) * N + cell(u);
}

int main(int argc, char **argv)
{
    const std::string base = argc > 1 ? argv[1] : "/tmp/lab_hall";
    {
        FILE *f = fopen((base + ".tri").c_str(), "rb");
        uint32_t n; if (!f || fread(&n, 4, 1, f) != 1) return 1;
        tris.resize(n);
        if (fread(tris.data(), sizeof(Tri), n, f) != n) return 1;
        fclose(f);
    }
    V3 mic = {24, 1.5f, 18};
    if (argc > 4) mic = {(float)atof(argv[2]), (float)atof(argv[3]), (float)atof(argv[4])};
    const float delta_rel = argc > 5 ? (float)atof(argv[5]) : 1e-4f;
    const float cpt = argc > 6 ? (float)atof(argv[6]) : 4.0f;
    g_exact = !(argc > 8 && atoi(argv[8]) == 0);
    const int T = (int)tris.size();
    V3 lo = {3e38f, 3e38f, 3e38f}, hi = {-3e38f, -3e38f, -3e38f};
    for (auto &t : tris) for (V3 p : {t.a, t.b, t.c}) {
        lo = {std::min(lo.x, p.x), std::min(lo.y, p.y), std::min(lo.z, p.z)}; hi = {std::max(hi.x, p.x), std::max(hi.y, p.y), std::max(hi.z, p.z)};
    }
    const float diag = lengthf(sub(hi, lo));
    const float delta = delta_rel * diag + 1e-6f, r_near = 40.0f * delta, wc = 0.5f * r_near;
    int N = (int)std::lround(std::sqrt(cpt * T / 6.0));
    N = std::min(512, std::max(8, N));
    printf("T %d  diag %.2f  delta %.4g  r_near %.3g  N %d (%d cells)  mic (%g %g %g)\n", T, diag, delta, r_near, N, 6 * N * N, mic.x, mic.y, mic.z);
    std::vector<std::vector<uint32_t>> cells(6 * N * N);
    std::vector<uint32_t> near;
    size_t entries = 0;
    for (int i = 0; i < T; ++i) {
        const Tri &t = tris[i];
        V3 p[3] = {sub(t.a, mic), sub(t.b, mic), sub(t.c, mic)};
        // distance from the mic to the triangle's bounding box (a lower bound of the distance to the triangle)
        float d2 = 0;
        for (int ax = 0; ax < 3; ++ax) {
            const float c[3] = {ax == 0 ? p[0].x : ax == 1 ? p[0].y : p[0].z, ax == 0 ? p[1].x : ax == 1 ? p[1].y : p[1].z, ax == 0 ? p[2].x : ax == 1 ? p[2].y : p[2].z};
            const float mn = std::min(c[0], std::min(c[1], c[2])), mx = std::max(c[0], std::max(c[1], c[2]));
            const float d = mn > 0 ? mn : (mx < 0 ? -mx : 0.0f);
            d2 += d * d;
        }
        if (d2 < r_near * r_near) { near.push_back(i); continue; }
        for (int f = 0; f < 6; ++f) {
            Rect r;
            if (!face_rect(p, f, wc, delta, N, r)) continue;
            for (int j = r.j0; j <= r.j1; ++j) for (int ii = r.i0; ii <= r.i1; ++ii) {
                if (cell_outside(r, ii, j, N)) continue;
                cells[(f * N + j) * N + ii].push_back(i); ++entries;
            }
        }
    }
    size_t mx = 0; for (auto &c : cells) mx = std::max(mx, c.size());
    printf("entries %zu (%.2f per triangle, %.2f per cell, max %zu), near list %zu\n", entries, (double)entries / T, (double)entries / cells.size(), mx, near.size());

    // queries: the dumped visibility queries + random surface points (vertices, edges, interiors)
    struct Q { V3 X; };
    std::vector<Q> qs;
    {
        FILE *f = fopen((base + ".qry").c_str(), "rb");
        uint32_t n; if (f && fread(&n, 4, 1, f) == 1) {
            std::vector<float> q(8 * (size_t)n);
            if (fread(q.data(), 4, q.size(), f) == q.size())
                for (uint32_t i = 0; i < n; ++i) if (q[8 * i + 7] == 1.0f) qs.push_back({{q[8 * i], q[8 * i + 1], q[8 * i + 2]}});
        }
        if (f) fclose(f);
    }
    const size_t nreal = qs.size();
    std::mt19937 rng(12345);
    std::uniform_real_distribution<float> U(0, 1);
    const int nrand = argc > 7 ? atoi(argv[7]) : 20000;
    
    for (int k = 0; k < nrand; ++k) {
        const Tri &t = tris[rng() % T];
        float a = U(rng), b = U(rng);
        if (a + b > 1) { a = 1 - a; b = 1 - b; }
        const int mode = k % 4;
        if (mode == 1) b = 0; else if (mode == 2) { a = 0; b = 0; } else if (mode == 3) { a = 1 - b; }
        V3 e0 = sub(t.b, t.a), e1 = sub(t.c, t.a);
        qs.push_back({{fmaf(e1.x, b, fmaf(e0.x, a, t.a.x)), fmaf(e1.y, b, fmaf(e0.y, a, t.a.y)), fmaf(e1.z, b, fmaf(e0.z, a, t.a.z))}});
    }
    size_t missing = 0, accepted = 0, blocked_q = 0;
    double cand_sum = 0, max_slop = 0, max_slop_rel = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : missing, accepted, blocked_q, cand_sum) reduction(max : max_slop, max_slop_rel)
    for (size_t qi = 0; qi < qs.size(); ++qi) {
        const V3 X = qs[qi].X;
        const V3 btp = sub(mic, X);
        const float mag = lengthf(btp);
        const V3 dir = normalizef(btp);
        const int cell = query_cell(sub(X, mic), N);
        const std::vector<uint32_t> &cand = cells[cell];
        cand_sum += cand.size() + near.size();
        bool blocked = false;
        for (int i = 0; i < T; ++i) {
            bool ok; const float tt = mt(tris[i], X, dir, ok);
            if (!(ok && tt > 1e-4f && tt <= mag)) continue;
            blocked = true; ++accepted;
            if (std::find(cand.begin(), cand.end(), (uint32_t)i) == cand.end() && std::find(near.begin(), near.end(), (uint32_t)i) == near.end()) {
                ++missing;
                if (missing < 5) fprintf(stderr, "MISSING tri %d for query %zu cell %d\n", i, qi, cell);
            }
            // how far outside the triangle is the true intersection with its plane? (double)
            const Tri &t = tris[i];
            double e0[3] = {(double)t.b.x - t.a.x, (double)t.b.y - t.a.y, (double)t.b.z - t.a.z}, e1[3] = {(double)t.c.x - t.a.x, (double)t.c.y - t.a.y, (double)t.c.z - t.a.z};
            double d[3] = {dir.x, dir.y, dir.z}, tv[3] = {(double)X.x - t.a.x, (double)X.y - t.a.y, (double)X.z - t.a.z};
            double p[3] = {d[1] * e1[2] - d[2] * e1[1], d[2] * e1[0] - d[0] * e1[2], d[0] * e1[1] - d[1] * e1[0]};
            double det = e0[0] * p[0] + e0[1] * p[1] + e0[2] * p[2];
            double u = (tv[0] * p[0] + tv[1] * p[1] + tv[2] * p[2]) / det;
            double q[3] = {tv[1] * e0[2] - tv[2] * e0[1], tv[2] * e0[0] - tv[0] * e0[2], tv[0] * e0[1] - tv[1] * e0[0]};
            double v = (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]) / det;
            double out = std::max(0.0, std::max(-u, std::max(-v, u + v - 1)));
            double edge = std::max(std::sqrt(e0[0] * e0[0] + e0[1] * e0[1] + e0[2] * e0[2]), std::sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]));
            max_slop = std::max(max_slop, out * edge);
            max_slop_rel = std::max(max_slop_rel, out);
        }
        blocked_q += blocked;
    }
    printf("%zu queries (%zu traced + %d random): candidates/query %.2f, blocked %zu, accepted hits %zu, MISSING %zu, max slop %.3g m (%.3g barycentric)\n",
           qs.size(), nreal, nrand, cand_sum / qs.size(), blocked_q, accepted, missing, max_slop, max_slop_rel);
    return missing != 0;
}
