// CPU lab for BVH builder / traversal-order decisions (no GPU needed): builds the candidate
// structures over a dumped scene and counts node visits / box tests / triangle tests over dumped,
// realistic queries (tools/bvh_lab/dump.py).  Development aid only -- not part of the product.
//   g++ -O2 -std=c++17 -o /tmp/lab tools/bvh_lab/lab.cpp && /tmp/lab /tmp/lab_hall
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <numeric>
#include <string>
#include <vector>

struct V3 { float x, y, z; };
static V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
struct Tri { V3 a, b, c; };
struct Box {
    V3 lo{3e38f, 3e38f, 3e38f}, hi{-3e38f, -3e38f, -3e38f};
    void grow(V3 p) { lo.x = std::min(lo.x, p.x); lo.y = std::min(lo.y, p.y); lo.z = std::min(lo.z, p.z); hi.x = std::max(hi.x, p.x); hi.y = std::max(hi.y, p.y); hi.z = std::max(hi.z, p.z); }
    void grow(const Box &b) { grow(b.lo); grow(b.hi); }
    float area() const { float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z; return dx < 0 ? 0.f : 2.f * (dx * dy + dy * dz + dz * dx); }
};
struct Query { float o[3], d[3], tmax, any; };

// binary BVH: node i children l[i], r[i]; child >= 0 inner, < 0 leaf: ~position in `order`
struct Bin {
    std::vector<int> l, r, first, last;   // [first,last] range in order
    std::vector<Box> box;
    std::vector<uint32_t> order;
    int root = 0;
};

static std::vector<Tri> tris;
static std::vector<Box> tbox;
static std::vector<V3> cen;

static uint64_t expand21(uint64_t v)
{
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

static std::vector<uint64_t> morton_codes(Box scene)
{
    std::vector<uint64_t> code(tris.size());
    const float S = 2097151.0f;
    V3 inv = {1.f / (scene.hi.x - scene.lo.x), 1.f / (scene.hi.y - scene.lo.y), 1.f / (scene.hi.z - scene.lo.z)};
    for (size_t i = 0; i < tris.size(); ++i) {
        uint64_t x = (uint64_t)std::min(std::max((cen[i].x - scene.lo.x) * inv.x * S, 0.f), S);
        uint64_t y = (uint64_t)std::min(std::max((cen[i].y - scene.lo.y) * inv.y * S, 0.f), S);
        uint64_t z = (uint64_t)std::min(std::max((cen[i].z - scene.lo.z) * inv.z * S, 0.f), S);
        code[i] = (expand21(x) << 2) | (expand21(y) << 1) | expand21(z);
    }
    return code;
}

static Box scene_box()
{
    Box b;
    for (auto &c : cen) b.grow(c);
    return b;
}

// ---- LBVH (same tree as Karras: split at the highest differing bit of (code, index)) --------------
static Bin build_lbvh()
{
    Bin B;
    const int n = (int)tris.size();
    auto code = morton_codes(scene_box());
    B.order.resize(n);
    std::iota(B.order.begin(), B.order.end(), 0u);
    std::stable_sort(B.order.begin(), B.order.end(), [&](uint32_t a, uint32_t b) { return code[a] < code[b]; });
    std::vector<uint64_t> sc(n);
    for (int i = 0; i < n; ++i) sc[i] = code[B.order[i]];
    auto delta = [&](int i, int j) -> int {
        if (sc[i] == sc[j]) return 64 + __builtin_clz((unsigned)i ^ (unsigned)j);
        return __builtin_clzll(sc[i] ^ sc[j]);
    };
    std::function<int(int, int)> rec = [&](int first, int last) -> int {
        if (first == last) return ~first;
        const int d = delta(first, last);
        int lo = first, hi = last;   // last index whose prefix with `first` is longer than d
        while (lo < hi) {
            int mid = (lo + hi + 1) / 2;
            if (delta(first, mid) > d) lo = mid; else hi = mid - 1;
        }
        const int id = (int)B.l.size();
        B.l.push_back(0); B.r.push_back(0); B.first.push_back(first); B.last.push_back(last); B.box.emplace_back();
        int L = rec(first, lo), R = rec(lo + 1, last);
        B.l[id] = L; B.r[id] = R;
        Box b;
        b.grow(L >= 0 ? B.box[L] : tbox[B.order[~L]]);
        b.grow(R >= 0 ? B.box[R] : tbox[B.order[~R]]);
        B.box[id] = b;
        return id;
    };
    B.root = rec(0, n - 1);
    return B;
}

// ---- binned SAH, top down ----------------------------------------------------------------------------
static Bin build_sah(int nbins = 16)
{
    Bin B;
    const int n = (int)tris.size();
    B.order.resize(n);
    std::iota(B.order.begin(), B.order.end(), 0u);
    std::function<int(int, int)> rec = [&](int first, int last) -> int {
        if (first == last) return ~first;
        Box cb, bb;
        for (int i = first; i <= last; ++i) { cb.grow(cen[B.order[i]]); bb.grow(tbox[B.order[i]]); }
        const int id = (int)B.l.size();
        B.l.push_back(0); B.r.push_back(0); B.first.push_back(first); B.last.push_back(last); B.box.push_back(bb);
        int best_axis = -1, best_split = 0;
        float best_cost = 3e38f;
        const float ext[3] = {cb.hi.x - cb.lo.x, cb.hi.y - cb.lo.y, cb.hi.z - cb.lo.z};
        const float lo[3] = {cb.lo.x, cb.lo.y, cb.lo.z};
        for (int ax = 0; ax < 3; ++ax) {
            if (ext[ax] <= 0) continue;
            std::vector<Box> bins(nbins);
            std::vector<int> cnt(nbins, 0);
            for (int i = first; i <= last; ++i) {
                const V3 c = cen[B.order[i]];
                const float v = ax == 0 ? c.x : ax == 1 ? c.y : c.z;
                int k = std::min(nbins - 1, (int)((v - lo[ax]) / ext[ax] * nbins));
                bins[k].grow(tbox[B.order[i]]); cnt[k]++;
            }
            std::vector<float> ra(nbins);
            std::vector<int> rc(nbins);
            Box acc; int c = 0;
            for (int k = nbins - 1; k > 0; --k) { acc.grow(bins[k]); c += cnt[k]; ra[k] = acc.area(); rc[k] = c; }
            acc = Box(); c = 0;
            for (int k = 0; k + 1 < nbins; ++k) {
                acc.grow(bins[k]); c += cnt[k];
                if (c == 0 || rc[k + 1] == 0) continue;
                const float cost = acc.area() * c + ra[k + 1] * rc[k + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = ax; best_split = k; }
            }
        }
        int mid;
        if (best_axis < 0) mid = (first + last) / 2;
        else {
            auto it = std::partition(B.order.begin() + first, B.order.begin() + last + 1, [&](uint32_t t) {
                const V3 c = cen[t];
                const float v = best_axis == 0 ? c.x : best_axis == 1 ? c.y : c.z;
                int k = std::min(nbins - 1, (int)((v - lo[best_axis]) / ext[best_axis] * nbins));
                return k <= best_split;
            });
            mid = (int)(it - B.order.begin()) - 1;
            if (mid < first || mid >= last) mid = (first + last) / 2;
        }
        int L = rec(first, mid), R = rec(mid + 1, last);
        B.l[id] = L; B.r[id] = R;
        return id;
    };
    B.root = rec(0, n - 1);
    return B;
}

// ---- PLOC (Meister & Bittner 2018): Morton order, nearest neighbour within a radius, merge mutual pairs
static Bin build_ploc(int radius)
{
    Bin B;
    const int n = (int)tris.size();
    auto code = morton_codes(scene_box());
    std::vector<uint32_t> ord(n);
    std::iota(ord.begin(), ord.end(), 0u);
    std::stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return code[a] < code[b]; });
    // the leaves are NOT contiguous ranges here: remap at the end by a DFS
    struct C { int ref; Box b; };
    std::vector<C> cur(n);
    for (int i = 0; i < n; ++i) cur[i] = {~i, tbox[ord[i]]};
    std::vector<int> l, r; std::vector<Box> nb;
    while (cur.size() > 1) {
        const int m = (int)cur.size();
        std::vector<int> nn(m);
        for (int i = 0; i < m; ++i) {
            float best = 3e38f; int bi = -1;
            for (int j = std::max(0, i - radius); j <= std::min(m - 1, i + radius); ++j) {
                if (j == i) continue;
                Box u = cur[i].b; u.grow(cur[j].b);
                const float a = u.area();
                if (a < best) { best = a; bi = j; }
            }
            nn[i] = bi;
        }
        std::vector<C> next;
        next.reserve(m);
        for (int i = 0; i < m; ++i) {
            const int j = nn[i];
            if (nn[j] == i) {
                if (i < j) {
                    Box u = cur[i].b; u.grow(cur[j].b);
                    const int id = (int)l.size();
                    l.push_back(cur[i].ref); r.push_back(cur[j].ref); nb.push_back(u);
                    next.push_back({id, u});
                }
            } else next.push_back(cur[i]);
        }
        cur.swap(next);
    }
    // DFS renumber so that every subtree covers a contiguous range of `order`
    const int root_old = cur[0].ref;
    B.order.clear();
    std::function<int(int)> dfs = [&](int ref) -> int {
        if (ref < 0) { B.order.push_back(ord[~ref]); return ~(int)(B.order.size() - 1); }
        const int id = (int)B.l.size();
        B.l.push_back(0); B.r.push_back(0); B.first.push_back(0); B.last.push_back(0); B.box.push_back(nb[ref]);
        const int f = (int)B.order.size();
        int L = dfs(l[ref]);
        int R = dfs(r[ref]);
        B.l[id] = L; B.r[id] = R; B.first[id] = f; B.last[id] = (int)B.order.size() - 1;
        return id;
    };
    B.root = dfs(root_old);
    return B;
}

static double sah_cost(const Bin &B)
{
    double c = 0;
    const double ra = B.box[B.root].area();
    for (size_t i = 0; i < B.l.size(); ++i) {
        c += 1.0 * B.box[i].area() / ra;
        if (B.l[i] < 0) c += 1.0 * tbox[B.order[~B.l[i]]].area() / ra;
        if (B.r[i] < 0) c += 1.0 * tbox[B.order[~B.r[i]]].area() / ra;
    }
    return c;
}

// ---- wide collapse -----------------------------------------------------------------------------------
struct Wide {
    int W;
    struct Node { int ref[8]; Box box[8]; int n; };
    std::vector<Node> nodes;
    std::vector<uint32_t> order;
};
// ref: >= 0 wide node; < 0: leaf ~(first | (cnt-1) << 28)
static Wide collapse(const Bin &B, int W, int leaf_max, bool by_area)
{
    Wide w; w.W = W; w.order = B.order;
    auto count = [&](int c) { return c < 0 ? 1 : B.last[c] - B.first[c] + 1; };
    auto cbox = [&](int c) { return c < 0 ? tbox[B.order[~c]] : B.box[c]; };
    std::function<int(int)> rec = [&](int b) -> int {
        const int id = (int)w.nodes.size();
        w.nodes.emplace_back();
        std::vector<int> kids = {B.l[b], B.r[b]};
        if (by_area) {
            while ((int)kids.size() < W) {
                int best = -1; float ba = -1;
                for (int k = 0; k < (int)kids.size(); ++k)
                    if (kids[k] >= 0 && count(kids[k]) > leaf_max && B.box[kids[k]].area() > ba) { ba = B.box[kids[k]].area(); best = k; }
                if (best < 0) break;
                const int c = kids[best];
                kids[best] = B.l[c];
                kids.push_back(B.r[c]);
            }
        } else {   // as the GPU build does today: one more level, uniformly (W = 4), two more for W = 8
            for (int lev = 1; (1 << (lev + 1)) <= W; ++lev) {
                std::vector<int> nk;
                for (int c : kids) {
                    if (c >= 0 && count(c) > leaf_max) { nk.push_back(B.l[c]); nk.push_back(B.r[c]); } else nk.push_back(c);
                }
                kids.swap(nk);
            }
        }
        Wide::Node nd; nd.n = (int)kids.size();
        for (int k = 0; k < nd.n; ++k) {
            const int c = kids[k];
            nd.box[k] = cbox(c);
            if (c < 0) nd.ref[k] = ~(int)((uint32_t)(~c));
            else if (count(c) <= leaf_max) nd.ref[k] = ~(int)((uint32_t)B.first[c] | ((uint32_t)(count(c) - 1) << 28));
            else nd.ref[k] = rec(c);
        }
        w.nodes[id] = nd;
        return id;
    };
    rec(B.root);
    return w;
}

static bool slab(const Box &b, const float *o, const float *inv, float tmax, float &tn)
{
    float tx0 = (b.lo.x - o[0]) * inv[0], tx1 = (b.hi.x - o[0]) * inv[0];
    float ty0 = (b.lo.y - o[1]) * inv[1], ty1 = (b.hi.y - o[1]) * inv[1];
    float tz0 = (b.lo.z - o[2]) * inv[2], tz1 = (b.hi.z - o[2]) * inv[2];
    float a = std::max(std::max(std::min(tx0, tx1), std::min(ty0, ty1)), std::max(std::min(tz0, tz1), 0.0f));
    float f = std::min(std::min(std::max(tx0, tx1), std::max(ty0, ty1)), std::min(std::max(tz0, tz1), tmax));
    tn = a;
    return a <= f * 1.00001f + 1e-6f;
}

static float mt(const Tri &t, const float *o, const float *d, bool &ok)
{
    V3 e0 = t.b - t.a, e1 = t.c - t.a, dir = {d[0], d[1], d[2]}, pos = {o[0], o[1], o[2]};
    V3 p = cross(dir, e1);
    float det = dot(e0, p);
    ok = false;
    if (det > -1e-4f && det < 1e-4f) return 0;
    float inv = 1.f / det;
    V3 tv = pos - t.a;
    float u = inv * dot(tv, p);
    if (u < 0 || u > 1) return 0;
    V3 q = cross(tv, e0);
    float v = inv * dot(dir, q);
    if (v < 0 || u + v > 1) return 0;
    ok = true;
    return inv * dot(e1, q);
}

struct Stats { double visits = 0, tri = 0, nq = 0, warp_max = 0, warp_sum = 0, nwarps = 0; };

static void run(const Wide &w, const std::vector<Query> &qs, int type, const char *name)
{
    Stats s;
    std::vector<int> per;
    for (const Query &q : qs) {
        if ((int)q.any != type) continue;
        float inv[3];
        for (int a = 0; a < 3; ++a) inv[a] = q.d[a] == 0 ? INFINITY : 1.f / q.d[a];
        float tbest = q.tmax; bool found = false;
        int stack[256], sp = 0, node = 0, visits = 0, ntri = 0;
        stack[sp++] = INT32_MIN;
        while (node != INT32_MIN) {
            if (node >= 0) {
                ++visits;
                const Wide::Node &nd = w.nodes[node];
                float t[8]; int r[8]; int m = 0;
                const float lim = tbest * 1.00001f + 1e-6f;
                for (int k = 0; k < nd.n; ++k) { float tn; if (slab(nd.box[k], q.o, inv, lim, tn)) { t[m] = tn; r[m] = nd.ref[k]; ++m; } }
                for (int i = 1; i < m; ++i) for (int j = i; j > 0 && t[j] < t[j - 1]; --j) { std::swap(t[j], t[j - 1]); std::swap(r[j], r[j - 1]); }
                for (int k = m - 1; k >= 1; --k) stack[sp++] = r[k];
                node = m ? r[0] : stack[--sp];
            } else {
                const uint32_t ref = ~(uint32_t)node;
                const uint32_t first = ref & 0x0fffffffu, cnt = (ref >> 28) + 1;
                for (uint32_t k = 0; k < cnt; ++k) {
                    bool ok; float tt = mt(tris[w.order[first + k]], q.o, q.d, ok);
                    ++ntri;
                    if (ok && tt > 1e-4f) {
                        if (type == 1) { if (tt <= q.tmax) found = true; }
                        else if (tt < tbest) tbest = tt;
                    }
                }
                node = (type == 1 && found) ? INT32_MIN : stack[--sp];
            }
        }
        s.visits += visits; s.tri += ntri; s.nq += 1;
        per.push_back(visits);
    }
    for (size_t i = 0; i + 32 <= per.size(); i += 32) {
        int mx = 0, sm = 0;
        for (int k = 0; k < 32; ++k) { mx = std::max(mx, per[i + k]); sm += per[i + k]; }
        s.warp_max += mx; s.warp_sum += sm / 32.0; s.nwarps += 1;
    }
    printf("  %-26s type %d: node visits/query %6.2f  box tests %6.1f  tri tests %5.2f  | warp: max %5.1f mean %5.1f (lane eff %.2f)\n", name, type,
           s.visits / s.nq, s.visits / s.nq * w.W, s.tri / s.nq, s.warp_max / s.nwarps, s.warp_sum / s.nwarps, s.warp_sum / s.warp_max);
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
    std::vector<Query> qs;
    {
        FILE *f = fopen((base + ".qry").c_str(), "rb");
        uint32_t n; if (!f || fread(&n, 4, 1, f) != 1) return 1;
        qs.resize(n);
        if (fread(qs.data(), sizeof(Query), n, f) != n) return 1;
        fclose(f);
    }
    tbox.resize(tris.size()); cen.resize(tris.size());
    for (size_t i = 0; i < tris.size(); ++i) {
        Box b; b.grow(tris[i].a); b.grow(tris[i].b); b.grow(tris[i].c);
        tbox[i] = b; cen[i] = {0.5f * (b.lo.x + b.hi.x), 0.5f * (b.lo.y + b.hi.y), 0.5f * (b.lo.z + b.hi.z)};
    }
    printf("%zu triangles, %zu queries\n", tris.size(), qs.size());
    struct Named { const char *name; Bin b; };
    std::vector<Named> bins;
    bins.push_back({"lbvh", build_lbvh()});
    bins.push_back({"sah16", build_sah(16)});
    bins.push_back({"ploc8", build_ploc(8)});
    bins.push_back({"ploc25", build_ploc(25)});
    for (auto &nb : bins) {
        printf("%s: SAH cost %.1f, %zu inner nodes\n", nb.name, sah_cost(nb.b), nb.b.l.size());
        for (int W : {4, 8}) for (int leaf : {4, 2, 1}) for (int by_area = 0; by_area < 2; ++by_area) {
            if (W == 8 && leaf != 4) continue;
            Wide w = collapse(nb.b, W, leaf, by_area != 0);
            char nm[64]; snprintf(nm, sizeof nm, "W%d leaf%d %s (%zu n)", W, leaf, by_area ? "area" : "unif", w.nodes.size());
            run(w, qs, 0, nm); run(w, qs, 1, nm);
        }
    }
    return 0;
}
