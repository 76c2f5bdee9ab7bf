/*
 * rayverb_oracle.c -- CPU restatement of the rayverb hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path may call, link or
 * import this file: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or the
 * reported CPU baseline.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * the upstream tree, reuk/parallel-reverb-raytracer).  The reference cannot be
 * built in this image (no OpenCL ICD/headers, no Assimp, no FFTW3), so this
 * restatement is pinned against the reference's own gtest golden vectors
 * (tests/raytrace_tests.h:35-47, tests/attenuation_tests.h:67-101,
 * tests/hrtf_tests.cpp:42-85) -- see tests/test_oracle_golden.py.  Everything
 * the reference does not test itself (volumes, times, image sources, dedupe,
 * flatten, process) is "parity unpinned" beyond this restatement; the biquad /
 * Linkwitz-Riley filters are additionally pinned against the reference's own
 * filters.cpp compiled into oracle/_ref (see oracle/Makefile).
 *
 * Canonical arithmetic.  OpenCL C leaves dot/cross/normalize precision and FP
 * contraction implementation-defined, so the reference's own bits are
 * implementation-defined.  This file pins ONE legal choice, written out with
 * explicit fmaf() and compiled with -ffp-contract=off so nothing else fuses:
 *
 *   dot(a,b)     = fma(a.z,b.z, fma(a.y,b.y, a.x*b.x))
 *   cross(a,b).x = fma(a.y,b.z, -(a.z*b.y))       (y, z cyclic)
 *   length(a)    = sqrt(dot(a,a))                   (IEEE sqrt)
 *   normalize(a) = l==0 ? 0 : a * (1/l)             (IEEE div; normalize(0)=0)
 *   a + b*s      = fma(b, s, a)   for ray points    (kernel.cpp:411,425,459)
 *   reflect      = fma(-(2n), dot(d,n), d)          (kernel.cpp:132)
 *
 * All of these are correctly-rounded IEEE-754 binary32 operations, so a GPU
 * kernel that performs the same operations in the same order produces the
 * same bits (hit indices, distances, positions, times).  Only pow/atan2 are
 * library functions whose last bits may differ between libm and CUDA.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EPSILON 0.0001f             /* kernel.cpp:11 */
#define NUM_IMAGE_SOURCE 10         /* clstructs.h:4 */
#define SPEED_OF_SOUND 340.0f       /* clstructs.h:5 */
#define RAY_GROUP_SIZE 4096         /* rayverb.h:199 */
#define NBANDS 8

/* ---- POD layouts, byte-identical to clstructs.h:13-58 ------------------- */
typedef struct { float s[4]; } cl_float3_t;                    /* 16 B */
typedef struct { float s[8]; } cl_float8_t;                    /* 32 B */
typedef struct { uint64_t surface, v0, v1, v2; } Triangle;     /* 32 B */
typedef struct { cl_float8_t specular, diffuse; } Surface;     /* 64 B */
typedef struct { cl_float8_t volume; cl_float3_t position; float time; float pad_[3]; } Impulse;          /* 64 B */
typedef struct { cl_float8_t volume; float time; float pad_[7]; } AttenuatedImpulse;                      /* 64 B */
typedef struct { cl_float3_t direction; float coefficient; float pad_[3]; } Speaker;                      /* 32 B */

typedef struct { float x, y, z; } v3;
typedef struct { v3 v0, v1, v2; } TriVerts;                    /* kernel.cpp:56-60 */

/* ---- canonical vector arithmetic ---------------------------------------- */
static inline v3 v_make(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 v_from(const cl_float3_t *p) { return v_make(p->s[0], p->s[1], p->s[2]); }
static inline v3 v_add(v3 a, v3 b) { return v_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v_sub(v3 a, v3 b) { return v_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v_scale(v3 a, float s) { return v_make(a.x * s, a.y * s, a.z * s); }
static inline v3 v_neg(v3 a) { return v_make(-a.x, -a.y, -a.z); }
static inline float v_dot(v3 a, v3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
static inline v3 v_cross(v3 a, v3 b)
{
    return v_make(fmaf(a.y, b.z, -(a.z * b.y)),
                  fmaf(a.z, b.x, -(a.x * b.z)),
                  fmaf(a.x, b.y, -(a.y * b.x)));
}
static inline float v_length(v3 a) { return sqrtf(v_dot(a, a)); }
static inline v3 v_normalize(v3 a)
{
    float l = v_length(a);
    if (l == 0.0f) return v_make(0.0f, 0.0f, 0.0f);   /* OpenCL: normalize(0) = 0 */
    return v_scale(a, 1.0f / l);
}
/* a + b*s, one rounding per component */
static inline v3 v_madd(v3 a, v3 b, float s) { return v_make(fmaf(b.x, s, a.x), fmaf(b.y, s, a.y), fmaf(b.z, s, a.z)); }

/* ---- kernel.cpp:63-88  triangle_vert_intersection ----------------------- */
static inline float triangle_vert_intersection(const TriVerts *v, v3 pos, v3 dir)
{
    v3 e0 = v_sub(v->v1, v->v0);
    v3 e1 = v_sub(v->v2, v->v0);
    v3 pvec = v_cross(dir, e1);
    float det = v_dot(e0, pvec);
    if (-EPSILON < det && det < EPSILON) return 0.0f;
    float invdet = 1.0f / det;
    v3 tvec = v_sub(pos, v->v0);
    float ucomp = invdet * v_dot(tvec, pvec);
    if (ucomp < 0.0f || 1.0f < ucomp) return 0.0f;
    v3 qvec = v_cross(tvec, e0);
    float vcomp = invdet * v_dot(dir, qvec);
    if (vcomp < 0.0f || 1.0f < vcomp + ucomp) return 0.0f;
    return invdet * v_dot(e1, qvec);
}

typedef struct {
    const Triangle *triangles; uint64_t ntriangles;
    const cl_float3_t *vertices; const Surface *surfaces;
} Scene;

static inline TriVerts scene_verts(const Scene *sc, uint64_t i)
{
    const Triangle *t = sc->triangles + i;
    TriVerts v = { v_from(sc->vertices + t->v0), v_from(sc->vertices + t->v1), v_from(sc->vertices + t->v2) };
    return v;
}

/* kernel.cpp:110-116 triangle_verts_normal */
static inline v3 triangle_verts_normal(const TriVerts *t)
{
    return v_normalize(v_cross(v_sub(t->v1, t->v0), v_sub(t->v2, t->v0)));
}

/* kernel.cpp:130-133 reflect: direction - (normal * 2 * dot(direction, normal)) */
static inline v3 reflect(v3 normal, v3 direction)
{
    float d = v_dot(direction, normal);
    v3 n2 = v_scale(normal, 2.0f);
    return v_make(fmaf(-n2.x, d, direction.x), fmaf(-n2.y, d, direction.y), fmaf(-n2.z, d, direction.z));
}

typedef struct { uint64_t primitive; float distance; int intersects; } Intersection;

/* kernel.cpp:167-192 ray_triangle_intersection (brute force; strict '<' =>
 * lowest triangle index wins exact ties) */
static inline Intersection ray_triangle_intersection(const Scene *sc, v3 pos, v3 dir, uint64_t *ntests)
{
    Intersection ret = {0, 0.0f, 0};
    for (uint64_t i = 0; i != sc->ntriangles; ++i) {
        TriVerts v = scene_verts(sc, i);
        float distance = triangle_vert_intersection(&v, pos, dir);
        if (distance > EPSILON && (!ret.intersects || distance < ret.distance)) {
            ret.primitive = i; ret.distance = distance; ret.intersects = 1;
        }
    }
    *ntests += sc->ntriangles;
    return ret;
}

/* kernel.cpp:195-214 attenuation_for_distance: pow(M_E, distance*AIR) * 1 */
static inline void attenuation_for_distance(float distance, const float *air, float *out)
{
    for (int b = 0; b < NBANDS; ++b) out[b] = powf((float)M_E, distance * air[b]);
}

/* kernel.cpp:217-221 mirror_point: *p += -n * dot(n, *p - t->v0) * 2 */
static inline void mirror_point(v3 *p, const TriVerts *t)
{
    v3 n = triangle_verts_normal(t);
    float d = v_dot(n, v_sub(*p, t->v0));
    v3 m = v_scale(v_scale(v_neg(n), d), 2.0f);
    *p = v_add(*p, m);
}
/* kernel.cpp:224-229 */
static inline void mirror_verts(TriVerts *in, const TriVerts *t)
{
    mirror_point(&in->v0, t); mirror_point(&in->v1, t); mirror_point(&in->v2, t);
}

/* kernel.cpp:274-296 point_intersection (visibility) */
static inline int point_intersection(const Scene *sc, v3 begin, v3 point, uint64_t *ntests)
{
    v3 btp = v_sub(point, begin);
    float mag = v_length(btp);
    v3 direction = v_normalize(btp);
    Intersection inter = ray_triangle_intersection(sc, begin, direction, ntests);
    return (!inter.intersects) || inter.distance > mag;
}

static inline void put3(cl_float3_t *d, v3 a) { d->s[0] = a.x; d->s[1] = a.y; d->s[2] = a.z; d->s[3] = 0.0f; }

/* kernel.cpp:243-265 add_image */
static inline void add_image(v3 mic_position, v3 mic_reflection, v3 source, Impulse *image_source,
                             uint64_t *image_source_index, size_t slot, const float *volume,
                             uint64_t object_index, const float *air)
{
    v3 init_diff = v_sub(source, mic_reflection);
    float init_dist = v_length(init_diff);
    float att[NBANDS];
    attenuation_for_distance(init_dist, air, att);
    Impulse *o = image_source + slot;
    for (int b = 0; b < NBANDS; ++b) o->volume.s[b] = volume[b] * att[b];
    put3(&o->position, v_add(mic_position, init_diff));
    o->time = (1.0f / SPEED_OF_SOUND) * init_dist;          /* kernel.cpp:14,262 */
    image_source_index[slot] = object_index;
}

/* kernel.cpp:304-503 kernel raytrace, one work-item.
 * impulses: row of nreflections slots (pre-zeroed by caller, rayverb.cpp:601-603)
 * image_source / image_source_index: row of NUM_IMAGE_SOURCE slots (pre-zeroed)
 * hit_tri (optional): per-bounce hit triangle index, 0xffffffff where no bounce happened
 * returns number of reflections executed.                                    */
static uint32_t raytrace_one(const Scene *sc, v3 direction, v3 position /* mic */, v3 source,
                             Impulse *impulses, Impulse *image_source, uint64_t *image_source_index,
                             uint64_t nreflections, const float *air, uint32_t *hit_tri, uint64_t *ntests)
{
    v3 ray_pos = source, ray_dir = direction;                  /* :327 */
    float distance = 0.0f;                                     /* :328 */
    float volume[NBANDS];
    for (int b = 0; b < NBANDS; ++b) volume[b] = 1.0f;         /* :329 */
    TriVerts prev_primitives[NUM_IMAGE_SOURCE - 1];            /* :332 */
    v3 mic_reflection = position;                              /* :333 */

    if (point_intersection(sc, source, mic_reflection, ntests))         /* :335-357 */
        add_image(position, mic_reflection, source, image_source, image_source_index, 0, volume, 0, air);

    uint32_t count = 0;
    for (uint64_t index = 0; index != nreflections; ++index) {          /* :359 */
        Intersection closest = ray_triangle_intersection(sc, ray_pos, ray_dir, ntests);   /* :363 */
        if (!closest.intersects) break;                                  /* :372 */
        const Triangle *triangle = sc->triangles + closest.primitive;
        if (hit_tri) hit_tri[index] = (uint32_t)closest.primitive;
        ++count;

        if (index < NUM_IMAGE_SOURCE - 1) {                              /* :379 */
            TriVerts current = scene_verts(sc, closest.primitive);
            for (unsigned k = 0; k != index; ++k) mirror_verts(&current, prev_primitives + k);   /* :387 */
            prev_primitives[index] = current;
            mirror_point(&mic_reflection, &current);                     /* :394 */
            v3 DIR = v_normalize(v_sub(mic_reflection, source));         /* :396 getDirection */
            int intersects = 1;
            v3 prevIntersection = source;
            for (uint64_t k = 0; k != index + 1 && intersects; ++k) {    /* :401 */
                float TO = triangle_vert_intersection(prev_primitives + k, source, DIR);
                if (TO <= EPSILON) { intersects = 0; break; }            /* :405 */
                v3 P = v_madd(source, DIR, TO);                          /* :411 */
                for (long l = (long)k - 1; l != -1; --l) mirror_point(&P, prev_primitives + l);
                v3 idir = v_normalize(v_sub(P, prevIntersection));       /* :417 */
                Intersection inter = ray_triangle_intersection(sc, prevIntersection, idir, ntests);
                v3 Q = v_madd(prevIntersection, idir, inter.distance);   /* :425 */
                intersects = inter.intersects
                    && (Q.x - EPSILON < P.x) && (Q.y - EPSILON < P.y) && (Q.z - EPSILON < P.z)
                    && (P.x < Q.x + EPSILON) && (P.y < Q.y + EPSILON) && (P.z < Q.z + EPSILON);  /* :426 */
                prevIntersection = P;
            }
            if (intersects) intersects = point_intersection(sc, prevIntersection, position, ntests);  /* :431 */
            if (intersects)                                              /* :442: pre-bounce volume (Q1) */
                add_image(position, mic_reflection, source, image_source, image_source_index,
                          index + 1, volume, closest.primitive + 1, air);
        }

        v3 intersection = v_madd(ray_pos, ray_dir, closest.distance);    /* :459 */
        float newDist = distance + closest.distance;                     /* :460 */
        const Surface *surf = sc->surfaces + triangle->surface;
        float newVol[NBANDS];
        for (int b = 0; b < NBANDS; ++b) newVol[b] = -volume[b] * surf->specular.s[b];   /* :461 */

        int vis = point_intersection(sc, intersection, position, ntests);                /* :463 */
        float DIST = vis ? newDist + v_length(v_sub(position, intersection)) : 0.0f;     /* :471 */
        TriVerts tv = scene_verts(sc, closest.primitive);
        v3 normal = triangle_verts_normal(&tv);
        float DIFF = fabsf(v_dot(normal, ray_dir));                                      /* :478 */
        Impulse *o = impulses + index;
        if (vis) {
            float att[NBANDS];
            attenuation_for_distance(DIST, air, att);
            for (int b = 0; b < NBANDS; ++b)
                o->volume.s[b] = ((newVol[b] * att[b]) * surf->diffuse.s[b]) * DIFF;     /* :481-485 */
        } else {
            for (int b = 0; b < NBANDS; ++b) o->volume.s[b] = 0.0f;
        }
        put3(&o->position, intersection);
        o->time = (1.0f / SPEED_OF_SOUND) * DIST;                                        /* :489 */

        ray_dir = reflect(normal, ray_dir);                                              /* :492-499 */
        ray_pos = intersection;
        distance = newDist;
        for (int b = 0; b < NBANDS; ++b) volume[b] = newVol[b];
    }
    return count;
}

/* AIR_COEFFICIENT as the host builds it: doubles narrowed to float
 * (rayverb.cpp:632-641) */
void orc_air_coefficient(float *air)
{
    const double c[NBANDS] = {0.001 * -0.1, 0.001 * -0.2, 0.001 * -0.5, 0.001 * -1.1,
                              0.001 * -2.7, 0.001 * -9.4, 0.001 * -29.0, 0.001 * -60.0};
    for (int b = 0; b < NBANDS; ++b) air[b] = (float)c[b];
}

/* Raytracer::raytrace device part (rayverb.cpp:587-651,678-683) over all rays.
 * diffuse      [nrays*nreflections]  ray-major, zero where no bounce
 * image        [nrays*10], image_idx [nrays*10]
 * hit_tri      [nrays*nreflections] or NULL; refl_count [nrays] or NULL
 * The 4096-wide tail work-items of a short last group (Q7) are defined to have
 * zero directions; such a ray only re-finds the direct path (key {0}), which
 * ray 0 has already supplied, so they are not simulated here.
 * Returns the number of ray-triangle tests executed.                        */
uint64_t orc_raytrace(const Triangle *triangles, uint64_t ntriangles, const cl_float3_t *vertices,
                      const Surface *surfaces, const float *mic, const float *source,
                      const cl_float3_t *directions, uint64_t nrays, uint64_t nreflections,
                      Impulse *diffuse, Impulse *image, uint64_t *image_idx,
                      uint32_t *hit_tri, uint32_t *refl_count)
{
    Scene sc = {triangles, ntriangles, vertices, surfaces};
    float air[NBANDS];
    orc_air_coefficient(air);
    v3 micv = v_make(mic[0], mic[1], mic[2]);
    v3 srcv = v_make(source[0], source[1], source[2]);
    memset(diffuse, 0, sizeof(Impulse) * nrays * nreflections);
    memset(image, 0, sizeof(Impulse) * nrays * NUM_IMAGE_SOURCE);
    memset(image_idx, 0, sizeof(uint64_t) * nrays * NUM_IMAGE_SOURCE);
    if (hit_tri) memset(hit_tri, 0xff, sizeof(uint32_t) * nrays * nreflections);
    uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : total)
    for (int64_t i = 0; i < (int64_t)nrays; ++i) {
        uint64_t nt = 0;
        uint32_t c = raytrace_one(&sc, v_from(directions + i), micv, srcv,
                                  diffuse + i * nreflections, image + i * NUM_IMAGE_SOURCE,
                                  image_idx + i * NUM_IMAGE_SOURCE, nreflections, air,
                                  hit_tri ? hit_tri + i * nreflections : NULL, &nt);
        if (refl_count) refl_count[i] = c;
        total += nt;
    }
    return total;
}

/* ---- image-source dedupe, rayverb.cpp:654-676 + getRawImages :692-706 ---- *
 * std::map<vector<ulong>, Impulse>: first writer (lowest ray) wins, output in
 * lexicographic key order (a proper prefix sorts before its extensions).      */
typedef struct { uint64_t key[NUM_IMAGE_SOURCE]; int len; uint64_t ray; const Impulse *imp; } Cand;

static int cand_cmp(const void *pa, const void *pb)
{
    const Cand *a = (const Cand *)pa, *b = (const Cand *)pb;
    int n = a->len < b->len ? a->len : b->len;
    for (int i = 0; i < n; ++i) {
        if (a->key[i] < b->key[i]) return -1;
        if (a->key[i] > b->key[i]) return 1;
    }
    if (a->len != b->len) return a->len < b->len ? -1 : 1;
    if (a->ray != b->ray) return a->ray < b->ray ? -1 : 1;
    return 0;
}

/* out must hold nrays*10 entries; out_keys (optional) nrays*10*10 uint64 (zero padded),
 * out_keylen (optional) nrays*10 ints.  Returns number of unique image impulses. */
uint64_t orc_image_dedupe(const Impulse *image, const uint64_t *image_idx, uint64_t nrays,
                          int remove_direct, Impulse *out, uint64_t *out_keys, int32_t *out_keylen)
{
    Cand *c = (Cand *)malloc(sizeof(Cand) * (size_t)(nrays * NUM_IMAGE_SOURCE + 1));
    size_t n = 0;
    for (uint64_t r = 0; r < nrays; ++r) {
        const uint64_t *idx = image_idx + r * NUM_IMAGE_SOURCE;
        for (int k = 1; k != NUM_IMAGE_SOURCE + 1; ++k) {
            if (k == 1 || idx[k - 1] != 0) {                              /* :667 */
                Cand *e = c + n++;
                memset(e->key, 0, sizeof e->key);
                for (int j = 0; j < k; ++j) e->key[j] = idx[j];
                e->len = k; e->ray = r; e->imp = image + r * NUM_IMAGE_SOURCE + (k - 1);
            }
        }
    }
    qsort(c, n, sizeof(Cand), cand_cmp);
    uint64_t m = 0;
    for (size_t i = 0; i < n; ++i) {
        if (i > 0 && c[i].len == c[i - 1].len && memcmp(c[i].key, c[i - 1].key, sizeof c[i].key) == 0) continue;
        if (remove_direct && c[i].len == 1 && c[i].key[0] == 0) continue;  /* temp.erase({0}) :696 */
        out[m] = *c[i].imp;
        if (out_keys) memcpy(out_keys + m * NUM_IMAGE_SOURCE, c[i].key, sizeof c[i].key);
        if (out_keylen) out_keylen[m] = c[i].len;
        ++m;
    }
    free(c);
    return m;
}

/* ---- kernel.cpp:506-535 speaker_attenuation + kernel attenuate ----------- *
 * Zero-volume impulses are skipped; the fresh output buffer is defined as
 * zeros (Q9).                                                                 */
void orc_attenuate_speaker(const Impulse *in, uint64_t n, const float *mic, const Speaker *speaker,
                           AttenuatedImpulse *out)
{
    v3 micv = v_make(mic[0], mic[1], mic[2]);
    v3 sdir = v_normalize(v_from(&speaker->direction));
    memset(out, 0, sizeof(AttenuatedImpulse) * n);
    for (uint64_t i = 0; i < n; ++i) {
        int any = 0;
        for (int b = 0; b < NBANDS; ++b) any |= (in[i].volume.s[b] != 0.0f);
        if (!any) continue;
        v3 dir = v_normalize(v_sub(v_from(&in[i].position), micv));          /* getDirection(mic, pos) */
        /* (1 - c) + c * dot(normalize(direction), normalize(speaker.direction)) */
        float att = (1.0f - speaker->coefficient) + speaker->coefficient * v_dot(v_normalize(dir), sdir);
        for (int b = 0; b < NBANDS; ++b) out[i].volume.s[b] = in[i].volume.s[b] * att;
        out[i].time = in[i].time;
    }
}

/* kernel.cpp:538-549 transform */
static inline v3 hrtf_transform(v3 pointing, v3 up, v3 d)
{
    v3 x = v_normalize(v_cross(up, pointing));
    v3 y = v_cross(pointing, x);
    v3 z = pointing;
    return v_make(v_dot(x, d), v_dot(y, d), v_dot(z, d));
}

/* kernel.cpp:563-625 hrtf_attenuation + kernel hrtf.  table = one ear,
 * [360][180] flattened a*180+e (rayverb.cpp:774-780).                        */
void orc_attenuate_hrtf(const Impulse *in, uint64_t n, const float *mic, const cl_float8_t *table,
                        const float *facing, const float *up, uint64_t channel, AttenuatedImpulse *out)
{
    v3 micv = v_make(mic[0], mic[1], mic[2]);
    v3 pointing = v_make(facing[0], facing[1], facing[2]);
    v3 upv = v_make(up[0], up[1], up[2]);
    const float WIDTH = 0.1f;
    v3 ear_pos = v_add(hrtf_transform(pointing, upv, v_make(channel == 0 ? -WIDTH : WIDTH, 0.0f, 0.0f)), micv);
    const float DEG = 57.295779513082320876798154814105f;                    /* degrees() */
    memset(out, 0, sizeof(AttenuatedImpulse) * n);
    for (uint64_t i = 0; i < n; ++i) {
        int any = 0;
        for (int b = 0; b < NBANDS; ++b) any |= (in[i].volume.s[b] != 0.0f);
        if (!any) continue;
        v3 pos = v_from(&in[i].position);
        v3 t = hrtf_transform(pointing, upv, v_normalize(v_sub(pos, micv)));
        float az = atan2f(t.x, t.z);                                         /* :552-555 */
        float el = atan2f(t.y, sqrtf(fmaf(t.z, t.z, t.x * t.x)));            /* :558-561 length(d.xz) */
        long a = (long)(DEG * az + 180.0f);                                  /* :578 */
        a %= 360;
        long e = (long)(DEG * el);
        e = 90 - e;
        const cl_float8_t *att = table + (a * 180 + e);
        float dist0 = v_length(v_sub(pos, micv));
        float dist1 = v_length(v_sub(pos, ear_pos));
        float diff = dist1 - dist0;
        for (int b = 0; b < NBANDS; ++b) out[i].volume.s[b] = in[i].volume.s[b] * att->s[b];
        out[i].time = in[i].time + diff * (1.0f / SPEED_OF_SOUND);            /* :622 */
    }
}

/* ---- rayverb.h:47-97 findPredelay / fixPredelay --------------------------- */
float orc_find_predelay(const AttenuatedImpulse *const *channels, const uint64_t *counts, uint64_t nchannels)
{
    /* nested accumulate: min over non-zero values, 0 if none */
    float acc_outer = 0.0f; int first_outer = 1;
    for (uint64_t c = 0; c < nchannels; ++c) {
        float acc = 0.0f;
        for (uint64_t i = 0; i < counts[c]; ++i) {
            float pd = channels[c][i].time;
            if (i == 0) { acc = pd; continue; }
            if (acc == 0) acc = pd; else if (pd == 0) { /* keep */ } else acc = acc < pd ? acc : pd;
        }
        if (first_outer) { acc_outer = acc; first_outer = 0; }
        else if (acc_outer == 0) acc_outer = acc;
        else if (acc == 0) { }
        else acc_outer = acc_outer < acc ? acc_outer : acc;
    }
    return acc_outer;
}

void orc_fix_predelay(AttenuatedImpulse *imp, uint64_t n, float seconds)
{
    for (uint64_t i = 0; i < n; ++i) imp[i].time = imp[i].time > seconds ? imp[i].time - seconds : 0.0f;
}

/* ---- rayverb.cpp:48-77 flattenImpulses (one channel) ---------------------- */
uint64_t orc_flatten_length(const AttenuatedImpulse *imp, uint64_t n, float samplerate)
{
    float maxtime = 0.0f;
    for (uint64_t i = 0; i < n; ++i) maxtime = maxtime > imp[i].time ? maxtime : imp[i].time;
    return (uint64_t)(roundf(maxtime * samplerate) + 1.0f);
}

/* out: [8][len] band-major */
void orc_flatten(const AttenuatedImpulse *imp, uint64_t n, float samplerate, float *out, uint64_t len)
{
    memset(out, 0, sizeof(float) * NBANDS * len);
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t sample = (uint64_t)roundf(imp[i].time * samplerate);      /* :69 */
        for (int j = 0; j < NBANDS; ++j) out[(uint64_t)j * len + sample] += imp[i].volume.s[j];
    }
}

/* ---- filters.cpp:156-168 Biquad::onepass / :185-191 twopass --------------- */
typedef struct { double b0, b1, b2, a1, a2; } Biquad;

static void biquad_onepass(const Biquad *q, float *data, uint64_t n)
{
    double z1 = 0, z2 = 0;
    for (uint64_t k = 0; k < n; ++k) {
        double x = data[k];
        double out = x * q->b0 + z1;
        z1 = x * q->b1 + z2 - q->a1 * out;
        z2 = x * q->b2 - q->a2 * out;
        data[k] = (float)out;
    }
}
static void reverse_f(float *d, uint64_t n)
{
    for (uint64_t i = 0, j = n ? n - 1 : 0; i < j; ++i, --j) { float t = d[i]; d[i] = d[j]; d[j] = t; }
}
static void biquad_twopass(const Biquad *q, float *data, uint64_t n)
{
    biquad_onepass(q, data, n); reverse_f(data, n); biquad_onepass(q, data, n); reverse_f(data, n);
}

/* filters.cpp:193-218 OnepassBandpassBiquad::setParams (RBJ band-pass) */
void orc_bandpass_biquad_params(float lo, float hi, float sr, double *coef)
{
    /* C++ overload resolution in the reference: sqrt(float) and log2(float)
     * pick the float overloads; everything else is double. */
    const double c = sqrtf(lo * hi);
    const double omega = 2 * M_PI * c / sr;
    const double cs = cos(omega);
    const double sn = sin(omega);
    const double bandwidth = log2f(hi / lo);
    const double Q = sn / (log(2) * bandwidth * omega);
    const double alpha = sn * sinh(1 / (2 * Q));
    const double a0 = 1 + alpha;
    const double nrm = 1 / a0;
    coef[0] = nrm * alpha; coef[1] = nrm * 0; coef[2] = nrm * -alpha;
    coef[3] = nrm * (-2 * cs); coef[4] = nrm * (1 - alpha);
}

/* filters.cpp:224-260 getC + LinkwitzRiley::setParams: coef[0..4] lopass, [5..9] hipass */
static double getC(double co, double sr) { const double wcT = M_PI * co / sr; return cos(wcT) / sin(wcT); }
void orc_linkwitz_riley_params(float l, float h, float s, double *coef)
{
    {
        const double c = getC(h, s);
        const double a0 = c * c + c * sqrt(2) + 1;
        coef[0] = 1 / a0; coef[1] = 2 / a0; coef[2] = 1 / a0;
        coef[3] = (-2 * (c * c - 1)) / a0; coef[4] = (c * c - c * sqrt(2) + 1) / a0;
    }
    {
        const double c = getC(l, s);
        const double a0 = c * c + c * sqrt(2) + 1;
        coef[5] = (c * c) / a0; coef[6] = (-2 * c * c) / a0; coef[7] = (c * c) / a0;
        coef[8] = (-2 * (c * c - 1)) / a0; coef[9] = (c * c - c * sqrt(2) + 1) / a0;
    }
}

/* ---- windowed-sinc crossover, filters.cpp:10-154,308-346 + filters.h:44-147 ------------------ *
 * The reference convolves through FFTW (r2c, pointwise product, c2r) with FFT_LENGTH =
 * 29 + inputLength - 1, so the circular convolution equals the linear one, and FFTW's c2r is
 * un-normalised: the result is FFT_LENGTH * (kernel (*) data).  FFTW is a third-party dependency
 * (FFTW 3.x, libfftw3f; not in the reference tree, not installed here); its published contract is
 * the unnormalised DFT, which is what is restated here as a direct convolution in double
 * precision.  Single-precision FFT rounding (~1e-7 of the signal norm) is the expected difference
 * from the reference's own output.                                                              */
#define SINC_KERNEL_LENGTH 29

/* filters.cpp:10-31 sincKernel<float>: centre tap 1, else sin(pi t)/(pi t), t = 2*cutoff*(i - (len-1)/2) */
static void sinc_kernel(double cutoff, int length, float *out)
{
    for (int i = 0; i != length; ++i) {
        if (i == (length - 1) / 2) out[i] = 1;
        else {
            double t = 2 * cutoff * (i - (length - 1) / 2.0);
            double pit = M_PI * t;
            out[i] = (float)(sin(pit) / pit);
        }
    }
}
/* filters.cpp:34-52 blackman<float> */
static void blackman(int length, float *out)
{
    const double a0 = 7938.0 / 18608.0, a1 = 9240.0 / 18608.0, a2 = 1430.0 / 18608.0;
    for (int i = 0; i != length; ++i) {
        const double offset = i / (length - 1.0);
        out[i] = (float)(a0 - a1 * cos(2 * M_PI * offset) + a2 * cos(4 * M_PI * offset));
    }
}
/* filters.cpp:56-70 lopassKernel: window * sinc (float product), normalised by the largest magnitude */
static void lopass_kernel(float sr, float cutoff, int length, float *out)
{
    float window[SINC_KERNEL_LENGTH];
    blackman(length, window);
    sinc_kernel(cutoff / sr, length, out);
    float m = 0;
    for (int i = 0; i != length; ++i) { out[i] = window[i] * out[i]; float a = fabsf(out[i]); m = m > a ? m : a; }
    const float f = (float)(1.0 / m);
    for (int i = 0; i != length; ++i) out[i] *= f;
}
/* filters.cpp:74-80 hipassKernel */
static void hipass_kernel(float sr, float cutoff, int length, float *out)
{
    lopass_kernel(sr, cutoff, length, out);
    for (int i = 0; i != length; ++i) out[i] = -out[i];
    out[(length - 1) / 2] += 1;
}
/* filters.cpp:114-124 bandpassKernel: FastConvolution(29).convolve(lop15, hip15) = 29 * (lop (*) hip) */
void orc_sinc_bandpass_kernel(float sr, float lo, float hi, float *kernel29)
{
    enum { HALF = 1 + SINC_KERNEL_LENGTH / 2 };
    float lop[HALF], hip[HALF];
    lopass_kernel(sr, hi, HALF, lop);
    hipass_kernel(sr, lo, HALF, hip);
    for (int k = 0; k < SINC_KERNEL_LENGTH; ++k) {
        double acc = 0;
        for (int j = 0; j < HALF; ++j) {
            int i = k - j;
            if (i >= 0 && i < HALF) acc += (double)lop[j] * (double)hip[i];
        }
        kernel29[k] = (float)(acc * SINC_KERNEL_LENGTH);
    }
}
/* filters.cpp:126-131 BandpassWindowedSinc::filter on one channel: in [8][len] -> out [8][fft_len],
 * fft_len = 29 + first_len - 1 where first_len is the length of the FIRST channel's first band
 * (filters.cpp:280-283: the convolver is sized once).  len must not exceed fft_len.             */
int orc_filter_channel_sinc(const float *in, uint64_t len, uint64_t first_len, float sr, float lo_cutoff, float *out)
{
    const float EDGES[9] = {lo_cutoff, 175, 350, 700, 1400, 2800, 5600, 11200, 20000};
    const uint64_t fft_len = SINC_KERNEL_LENGTH + first_len - 1;
    if (len > fft_len) return -2;
    for (int b = 0; b < NBANDS; ++b) {
        float k[SINC_KERNEL_LENGTH];
        orc_sinc_bandpass_kernel(sr, EDGES[b], EDGES[b + 1], k);
        const float *x = in + (uint64_t)b * len;
        float *y = out + (uint64_t)b * fft_len;
        for (uint64_t n = 0; n < fft_len; ++n) {
            double acc = 0;
            for (int j = 0; j < SINC_KERNEL_LENGTH; ++j) {
                if (n >= (uint64_t)j && n - j < len) acc += (double)k[j] * (double)x[n - j];
            }
            y[n] = (float)(acc * (double)fft_len);
        }
    }
    return 0;
}

/* FilterType, filters.h:196-201 */
enum { FILTER_TYPE_WINDOWED_SINC = 0, FILTER_TYPE_BIQUAD_ONEPASS = 1, FILTER_TYPE_BIQUAD_TWOPASS = 2, FILTER_TYPE_LINKWITZ_RILEY = 3 };

/* filters.cpp:268-306 RayverbFiltering::filter for one channel [8][len].
 * Returns 0, or -1 for the windowed-sinc type (needs FFTW; "next" row).     */
int orc_filter_channel(int ft, float *data, uint64_t len, float sr, float lo_cutoff)
{
    const float EDGES[9] = {lo_cutoff, 175, 350, 700, 1400, 2800, 5600, 11200, 20000};
    for (int i = 0; i < NBANDS; ++i) {
        float *band = data + (uint64_t)i * len;
        if (ft == FILTER_TYPE_BIQUAD_ONEPASS || ft == FILTER_TYPE_BIQUAD_TWOPASS) {
            Biquad q; orc_bandpass_biquad_params(EDGES[i], EDGES[i + 1], sr, &q.b0);
            if (ft == FILTER_TYPE_BIQUAD_ONEPASS) biquad_onepass(&q, band, len); else biquad_twopass(&q, band, len);
        } else if (ft == FILTER_TYPE_LINKWITZ_RILEY) {
            double c[10]; orc_linkwitz_riley_params(EDGES[i], EDGES[i + 1], sr, c);
            Biquad lp = {c[0], c[1], c[2], c[3], c[4]}, hp = {c[5], c[6], c[7], c[8], c[9]};
            biquad_twopass(&lp, band, len);                                   /* filters.cpp:262-266 */
            biquad_twopass(&hp, band, len);
        } else {
            return -1;
        }
    }
    return 0;
}

/* rayverb.cpp:80-92 mixdown: sum bands, in band order, float accumulation */
void orc_mixdown(const float *data, uint64_t len, float *out)
{
    for (uint64_t k = 0; k < len; ++k) out[k] = 0.0f;
    for (int i = 0; i < NBANDS; ++i)
        for (uint64_t k = 0; k < len; ++k) out[k] = out[k] + data[(uint64_t)i * len + k];
}

/* generic_functions.h:8-62: max_amp (float accumulate of fabs), normalize = mul(1.0/max) */
float orc_max_amp(const float *const *channels, const uint64_t *lens, uint64_t nch)
{
    float m = 0.0f;
    for (uint64_t c = 0; c < nch; ++c)
        for (uint64_t k = 0; k < lens[c]; ++k) { float a = fabsf(channels[c][k]); m = m > a ? m : a; }
    return m;
}
void orc_mul(float *d, uint64_t n, float f) { for (uint64_t k = 0; k < n; ++k) d[k] *= f; }

/* rayverb.cpp:96-122 trimTail: len = max over channels of (index of last
 * |x| >= minVol), 0 if none; every channel is resized to len (Q12: the last
 * above-threshold sample itself is dropped).                                  */
int64_t orc_trim_length(const float *const *channels, const uint64_t *lens, uint64_t nch, float minVol)
{
    long len = 0;
    for (uint64_t c = 0; c < nch; ++c) {
        long last = -1;  /* find_if(rbegin..).base() - begin - 1 */
        for (int64_t k = (int64_t)lens[c] - 1; k >= 0; --k)
            if (fabsf(channels[c][k]) >= minVol) { last = (long)k; break; }
        if (last > len) len = last;
    }
    return len;
}

/* rayverb.cpp:125-149 process, for channels already flattened to [8][len_c].
 * in: nch pointers to [8][lens[c]]; out: nch pointers to [lens[c]] mixed
 * channels (caller allocates); out_len = common trimmed length if
 * do_trim_tail else -1 (each channel keeps lens[c]).                          */
int orc_process(int ft, float *const *data, const uint64_t *lens, uint64_t nch, float sr, int do_normalize,
                float lo_cutoff, int do_trim_tail, float volume_scale, float *const *out, int64_t *out_len)
{
    for (uint64_t c = 0; c < nch; ++c) {
        int rc = orc_filter_channel(ft, data[c], lens[c], sr, lo_cutoff);
        if (rc) return rc;
        orc_mixdown(data[c], lens[c], out[c]);
    }
    if (do_normalize) {
        float m = orc_max_amp((const float *const *)out, lens, nch);
        float f = (float)(1.0 / m);                                          /* mul(ret, 1.0 / max_amp) */
        for (uint64_t c = 0; c < nch; ++c) orc_mul(out[c], lens[c], f);
    }
    if (volume_scale != 1) for (uint64_t c = 0; c < nch; ++c) orc_mul(out[c], lens[c], volume_scale);
    *out_len = -1;
    if (do_trim_tail) *out_len = orc_trim_length((const float *const *)out, lens, nch, 0.00001f);
    return 0;
}

/* helpers.cpp:63-67 spherePoint */
void orc_sphere_point(float z, float theta, float *out)
{
    const float ztemp = sqrtf(1 - z * z);
    out[0] = ztemp * cosf(theta); out[1] = ztemp * sinf(theta); out[2] = z; out[3] = 0;
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU baseline legs ask for the host's cores */
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
