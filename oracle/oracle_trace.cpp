/*
 * oracle_trace.cpp -- CPU ORACLE (test infrastructure, not the product).
 * Restates OR/devicePrograms.cu:62-254; pinned against that file compiled where it lies (arv2_oracle.h "PIN STATUS").
 *
 * Build: g++ -O3 -std=c++17 -ffp-contract=off -mavx2 -mfma  (see oracle/Makefile; results do not depend on -O / -march).
 * Every a*b+c that must be fused is an explicit fmaf(); everything else is
 * evaluated exactly as written (no contraction), IEEE sqrt and division.
 *
 * Where the reference delegates arithmetic to a black box the oracle fixes one
 * concrete recipe (DESIGN.md "arithmetic contract"):
 *   - optixTrace closest hit (devicePrograms.cu:240-251): Moller-Trumbore on
 *     (P1, P2-P1, P3-P1), two-sided, hit iff t > 0 and t < 1e20; closest = min
 *     t, ties broken by the lower global triangle id.
 *   - curand_init(clock64(), tid)/curand_uniform (devicePrograms.cu:216-220):
 *     Philox4x32-10 keyed by an explicit 64-bit seed, counter = (ray id, bounce,
 *     purpose); same distribution theta = 2*pi*u1, phi = acos(2*u2-1).
 */
#include "arv2_oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct V3 { float x, y, z; };

inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float dot(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
inline V3 cross(V3 a, V3 b)
{
    return {fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))};
}

/* ------------------------------------------------------------------ RNG -- */
/* Philox4x32-10 (Salmon et al. 2011); counter (ray_lo, ray_hi, bounce, purpose),
 * key (seed_lo, seed_hi). */
inline void philox(uint64_t seed, uint64_t ray, uint32_t bounce, uint32_t purpose, uint32_t out[4])
{
    uint32_t c0 = (uint32_t)ray, c1 = (uint32_t)(ray >> 32), c2 = bounce, c3 = purpose;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* (cos, sin) of 2*pi*(k+0.5)/2^32 in fp64 from integer range reduction and two
 * Taylor polynomials evaluated with fma only, so that any IEEE machine gets the
 * same bits.  Coefficients are the doubles nearest (-1)^i/(2i+1)!, (-1)^i/(2i)!. */
inline void sincos_turn(uint32_t k, double* c_out, double* s_out)
{
    const uint32_t q = k >> 30;
    double f = ((double)(k & 0x3FFFFFFFu) + 0.5) * 0x1p-30; /* (0,1) quarter-turn fraction */
    const bool swap = f > 0.5;
    if (swap) f = 1.0 - f;
    const double b = f * 0x1.921fb54442d18p+0; /* pi/2 */
    const double b2 = b * b;
    double ps = 0x1.952c77030ad4ap-49;         /*  1/17! */
    ps = fma(ps, b2, -0x1.ae7f3e733b81fp-41);  /* -1/15! */
    ps = fma(ps, b2, 0x1.6124613a86d09p-33);   /*  1/13! */
    ps = fma(ps, b2, -0x1.ae64567f544e4p-26);  /* -1/11! */
    ps = fma(ps, b2, 0x1.71de3a556c734p-19);   /*  1/9!  */
    ps = fma(ps, b2, -0x1.a01a01a01a01ap-13);  /* -1/7!  */
    ps = fma(ps, b2, 0x1.1111111111111p-7);    /*  1/5!  */
    ps = fma(ps, b2, -0x1.5555555555555p-3);   /* -1/3!  */
    double s = fma(b * b2, ps, b);
    double pc = 0x1.ae7f3e733b81fp-45;         /*  1/16! */
    pc = fma(pc, b2, -0x1.93974a8c07c9dp-37);  /* -1/14! */
    pc = fma(pc, b2, 0x1.1eed8eff8d898p-29);   /*  1/12! */
    pc = fma(pc, b2, -0x1.27e4fb7789f5cp-22);  /* -1/10! */
    pc = fma(pc, b2, 0x1.a01a01a01a01ap-16);   /*  1/8!  */
    pc = fma(pc, b2, -0x1.6c16c16c16c17p-10);  /* -1/6!  */
    pc = fma(pc, b2, 0x1.5555555555555p-5);    /*  1/4!  */
    pc = fma(pc, b2, -0x1p-1);                 /* -1/2!  */
    double c = fma(pc, b2, 1.0);
    if (swap) { double tmp = s; s = c; c = tmp; }
    switch (q) {
    case 0: *c_out = c;  *s_out = s;  break;
    case 1: *c_out = -s; *s_out = c;  break;
    case 2: *c_out = -c; *s_out = -s; break;
    default: *c_out = s; *s_out = -c; break;
    }
}

/* devicePrograms.cu:219-224: theta = 2*pi*u1, phi = acos(2*u2-1),
 * (sin(phi)cos(theta), sin(phi)sin(theta), cos(phi)) in double, narrowed. */
inline V3 emit_direction(uint64_t seed, uint64_t ray)
{
    uint32_t r[4];
    philox(seed, ray, 0u, 0u, r);
    double ct, st;
    sincos_turn(r[0], &ct, &st);
    const double u2 = (double)((r[1] >> 8) + 1u) * 0x1p-24; /* (0,1] like curand_uniform */
    const double z = 2.0 * u2 - 1.0;                        /* cos(phi) */
    const double sp = sqrt(fma(-z, z, 1.0));                /* sin(phi) */
    return {(float)(sp * ct), (float)(sp * st), (float)z};
}

/* ------------------------------------------------------- intersection --- */
struct Tri { V3 p1, e1, e2; };
struct Hit { float t, u, v; int64_t id; };

/* Two-sided Moller-Trumbore in scaled form; division only for accepted hits.
 * NaNs fall out as misses (all predicates are positive-form). */
inline bool tri_test(const Tri& tr, V3 org, V3 dir, float* t, float* u, float* v)
{
    const V3 pvec = cross(dir, tr.e2);
    float det = dot(tr.e1, pvec);
    if (det == 0.0f) return false;
    const V3 tvec = sub(org, tr.p1);
    float U = dot(tvec, pvec);
    const V3 qvec = cross(tvec, tr.e1);
    float V = dot(dir, qvec);
    float T = dot(tr.e2, qvec);
    if (det < 0.0f) { det = -det; U = -U; V = -V; T = -T; }
    if (!(U >= 0.0f && V >= 0.0f && U + V <= det)) return false;
    if (!(T > 0.0f)) return false;
    const float tt = T / det;
    if (!(tt < 1e20f)) return false;
    *t = tt; *u = U / det; *v = V / det;
    return true;
}

inline void consider(Hit& best, float t, float u, float v, int64_t id)
{
    if (t < best.t || (t == best.t && id < best.id)) { best.t = t; best.u = u; best.v = v; best.id = id; }
}

/* ---------------------------------------------------------------- BVH ---- */
/* The oracle's own accelerator: binned-SAH binary BVH over triangle centroids, leaves of
 * <= 4, front-to-back descent by the sign of the ray on the split axis, boxes padded so that
 * the float slab test never rejects a box whose triangle the exact test would accept.
 * Written independently of the product's builders; it cannot change results, only speed. */
struct Node { float lo[3], hi[3]; int32_t left, right, first, count, axis; };

struct Bvh {
    std::vector<Node> nodes;
    std::vector<int64_t> order;
};

struct TriBox { float lo[3], hi[3], c[3]; };

void build_rec(Bvh& b, const std::vector<TriBox>& tb, float pad, int node, int64_t first, int64_t count)
{
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = first; i < first + count; ++i) {
        const TriBox& t = tb[b.order[i]];
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], t.lo[a]); hi[a] = std::max(hi[a], t.hi[a]);
            clo[a] = std::min(clo[a], t.c[a]); chi[a] = std::max(chi[a], t.c[a]);
        }
    }
    Node n;
    for (int a = 0; a < 3; ++a) { n.lo[a] = lo[a] - pad; n.hi[a] = hi[a] + pad; }
    n.left = n.right = -1; n.first = (int32_t)first; n.count = (int32_t)count; n.axis = 0;
    if (count <= 4) { b.nodes[node] = n; return; }

    /* binned surface-area heuristic, 12 bins per axis */
    const int NB = 12;
    float best = INFINITY; int best_axis = -1, best_bin = -1;
    auto half_area = [](const float* l, const float* h) {
        const float x = h[0] - l[0], y = h[1] - l[1], z = h[2] - l[2];
        return (x >= 0 && y >= 0 && z >= 0) ? x * y + y * z + z * x : 0.f;
    };
    for (int a = 0; a < 3; ++a) {
        const float ext = chi[a] - clo[a];
        if (!(ext > 0.f)) continue;
        float blo[NB][3], bhi[NB][3]; int cnt[NB];
        for (int k = 0; k < NB; ++k) { cnt[k] = 0; for (int d = 0; d < 3; ++d) { blo[k][d] = INFINITY; bhi[k][d] = -INFINITY; } }
        for (int64_t i = first; i < first + count; ++i) {
            const TriBox& t = tb[b.order[i]];
            int k = std::min(NB - 1, std::max(0, (int)((t.c[a] - clo[a]) / ext * NB)));
            cnt[k]++;
            for (int d = 0; d < 3; ++d) { blo[k][d] = std::min(blo[k][d], t.lo[d]); bhi[k][d] = std::max(bhi[k][d], t.hi[d]); }
        }
        float rarea[NB]; int rcnt[NB];
        float al[3] = {INFINITY, INFINITY, INFINITY}, ah[3] = {-INFINITY, -INFINITY, -INFINITY}; int c = 0;
        for (int k = NB - 1; k > 0; --k) {
            for (int d = 0; d < 3; ++d) { al[d] = std::min(al[d], blo[k][d]); ah[d] = std::max(ah[d], bhi[k][d]); }
            c += cnt[k]; rarea[k] = half_area(al, ah); rcnt[k] = c;
        }
        for (int d = 0; d < 3; ++d) { al[d] = INFINITY; ah[d] = -INFINITY; }
        c = 0;
        for (int k = 0; k < NB - 1; ++k) {
            for (int d = 0; d < 3; ++d) { al[d] = std::min(al[d], blo[k][d]); ah[d] = std::max(ah[d], bhi[k][d]); }
            c += cnt[k];
            if (c == 0 || rcnt[k + 1] == 0) continue;
            const float cost = half_area(al, ah) * c + rarea[k + 1] * rcnt[k + 1];
            if (cost < best) { best = cost; best_axis = a; best_bin = k; }
        }
    }
    int64_t mid;
    if (best_axis >= 0) {
        const int a = best_axis; const float ext = chi[a] - clo[a], base = clo[a];
        auto it = std::partition(b.order.begin() + first, b.order.begin() + first + count, [&](int64_t id) {
            return std::min(NB - 1, std::max(0, (int)((tb[id].c[a] - base) / ext * NB))) <= best_bin;
        });
        mid = it - b.order.begin();
        n.axis = a;
    } else {
        mid = first + count / 2;
    }
    if (mid == first || mid == first + count) mid = first + count / 2;
    n.count = 0;
    n.left = (int32_t)b.nodes.size(); b.nodes.emplace_back();
    n.right = (int32_t)b.nodes.size(); b.nodes.emplace_back();
    b.nodes[node] = n;
    build_rec(b, tb, pad, n.left, first, mid - first);
    build_rec(b, tb, pad, n.right, mid, first + count - mid);
}

Bvh build_bvh(const std::vector<Tri>& tris)
{
    Bvh b;
    const int64_t T = (int64_t)tris.size();
    b.order.resize(T);
    std::vector<TriBox> tb(T);
    float ext = 0.f;
    for (int64_t i = 0; i < T; ++i) {
        b.order[i] = i;
        const Tri& t = tris[i];
        const float v[3][3] = {{t.p1.x, t.p1.y, t.p1.z},
                               {t.p1.x + t.e1.x, t.p1.y + t.e1.y, t.p1.z + t.e1.z},
                               {t.p1.x + t.e2.x, t.p1.y + t.e2.y, t.p1.z + t.e2.z}};
        for (int a = 0; a < 3; ++a) {
            /* P1 + e is not exactly P2/P3 in float: widen by the rounding of the sum */
            const float m = std::max(std::fabs(v[0][a]), std::max(std::fabs(v[1][a]), std::fabs(v[2][a]))) * 2e-7f;
            tb[i].lo[a] = std::min(v[0][a], std::min(v[1][a], v[2][a])) - m;
            tb[i].hi[a] = std::max(v[0][a], std::max(v[1][a], v[2][a])) + m;
            tb[i].c[a] = 0.5f * (tb[i].lo[a] + tb[i].hi[a]);
            ext = std::max(ext, std::max(std::fabs(tb[i].lo[a]), std::fabs(tb[i].hi[a])));
        }
    }
    const float pad = ext * 1e-5f + 1e-6f;
    b.nodes.reserve((size_t)T + 2);
    b.nodes.emplace_back();
    if (T > 0) build_rec(b, tb, pad, 0, 0, T);
    return b;
}

inline bool slab(const Node& n, V3 o, V3 inv, float tmax)
{
    float t0 = 0.f, t1 = tmax;
    const float oo[3] = {o.x, o.y, o.z}, ii[3] = {inv.x, inv.y, inv.z};
    for (int a = 0; a < 3; ++a) {
        float ta = (n.lo[a] - oo[a]) * ii[a], tb = (n.hi[a] - oo[a]) * ii[a];
        if (ta > tb) std::swap(ta, tb);
        /* NaN (0*inf) never tightens the interval */
        if (ta > t0) t0 = ta;
        if (tb < t1) t1 = tb;
    }
    return t0 <= t1 * 1.0000005f;
}

Hit closest_bvh(const Bvh& b, const std::vector<Tri>& tris, V3 org, V3 dir)
{
    Hit best{1e20f, 0.f, 0.f, -1};
    if (tris.empty()) return best;
    const V3 inv = {1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z};
    const float dd[3] = {dir.x, dir.y, dir.z};
    int stack[256];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const Node& n = b.nodes[stack[--sp]];
        if (!slab(n, org, inv, best.t)) continue;
        if (n.left < 0) {
            for (int i = 0; i < n.count; ++i) {
                const int64_t id = b.order[n.first + i];
                float t, u, v;
                if (tri_test(tris[id], org, dir, &t, &u, &v)) consider(best, t, u, v, id);
            }
        } else if (sp + 2 <= 256) {
            /* near child first: the split ordered `left` below `right` on n.axis */
            if (dd[n.axis] >= 0.f) { stack[sp++] = n.right; stack[sp++] = n.left; }
            else { stack[sp++] = n.left; stack[sp++] = n.right; }
        }
    }
    return best;
}

Hit closest_brute(const std::vector<Tri>& tris, V3 org, V3 dir)
{
    Hit best{1e20f, 0.f, 0.f, -1};
    for (int64_t id = 0; id < (int64_t)tris.size(); ++id) {
        float t, u, v;
        if (tri_test(tris[id], org, dir, &t, &u, &v)) consider(best, t, u, v, id);
    }
    return best;
}

std::vector<Tri> make_tris(const float* tv, int64_t T)
{
    std::vector<Tri> tris(T);
    for (int64_t i = 0; i < T; ++i) {
        const float* p = tv + 9 * i;
        const V3 p1{p[0], p[1], p[2]}, p2{p[3], p[4], p[5]}, p3{p[6], p[7], p[8]};
        tris[i] = {p1, sub(p2, p1), sub(p3, p1)};
    }
    return tris;
}

/* Lambert bounce (new-build extension; never taken when scattering == 0).
 * Cosine-weighted direction about the normal that faces the incoming ray, in
 * the branchless orthonormal basis of Duff et al. 2017. */
inline V3 lambert_direction(const uint32_t r[4], V3 dir, V3 ng)
{
    V3 n = ng;
    if (dot(dir, ng) > 0.0f) n = {-ng.x, -ng.y, -ng.z};
    double cp, sp;
    sincos_turn(r[2], &cp, &sp);
    const double u = ((double)(r[1] >> 8) + 0.5) * 0x1p-24; /* (0,1) */
    const double sr = sqrt(u), cz = sqrt(1.0 - u);
    const float lx = (float)(sr * cp), ly = (float)(sr * sp), lz = (float)cz;
    const float sg = copysignf(1.0f, n.z);
    const float a = -1.0f / (sg + n.z);
    const float b = n.x * n.y * a;
    const V3 t1{fmaf(sg * n.x, n.x * a, 1.0f), sg * b, -sg * n.x};
    const V3 t2{b, fmaf(n.y, n.y * a, sg), -n.y};
    return {fmaf(lz, n.x, fmaf(ly, t2.x, lx * t1.x)), fmaf(lz, n.y, fmaf(ly, t2.y, lx * t1.y)),
            fmaf(lz, n.z, fmaf(ly, t2.z, lx * t1.z))};
}

constexpr int MAX_BANDS = 8;

struct Scene {
    std::vector<Tri> tris;
    std::vector<V3> p2, p3;
    const int32_t* mat;
    const float* absorption;
    const float* scattering;
    Bvh bvh;
    bool use_bvh;
};

struct RayOut { int32_t bin, ear, nseg; float energy[MAX_BANDS]; };

/* struct PRD (OR/PRD.h:5-14) with one remaining_factor per band */
struct PathState { float energy[MAX_BANDS]; float distance; V3 prev, dir; int depth; };

/* the atomicAdds of one closest-hit invocation (<= 2 per band): ear 0 = ir_left, 1 = ir_right */
struct Deposits { int n; int ear[2]; int idx[2]; float val[2][MAX_BANDS]; };

/* __closesthit__radiance (devicePrograms.cu:62-180) for the hit (u, v) on triangle (p1, p2, p3) of material m
 * (>= 0 wall, -1 receiver_left, -2 receiver_right): updates the path state, lists what it deposits.
 * keep = the wall's absorption row [nb]; sc = its scattering coefficient (0 = reference). */
inline void shade_hit(const oracle_params& P, int nb, V3 p1, V3 p2, V3 p3, int m, const float* absorption, float sc,
                      uint64_t ray, V3 center, float u, float v, int delay, PathState& st, Deposits& dep)
{
    dep.n = 0;
    /* :75-77  Ng = normalize(cross(P2-P1, P3-P1)) */
    const V3 nc = cross(sub(p2, p1), sub(p3, p1));
    const float ninv = 1.0f / sqrtf(dot(nc, nc));
    const V3 ng{nc.x * ninv, nc.y * ninv, nc.z * ninv};
    /* :81  P = (1-u-v)P1 + uP2 + vP3 */
    const float w = (1.0f - u) - v;
    const V3 pt{fmaf(v, p3.x, fmaf(u, p2.x, w * p1.x)), fmaf(v, p3.y, fmaf(u, p2.y, w * p1.y)),
                fmaf(v, p3.z, fmaf(u, p2.z, w * p1.z))};
    /* :83  distance += |P - prev_position| */
    const V3 dp = sub(pt, st.prev);
    st.distance = st.distance + sqrtf(dot(dp, dp));
    const V3 dir = st.dir;

    if (m < 0) {
        /* :91-122  chord of the unit ball around sphere_center along the ray */
        const float dinv = 1.0f / sqrtf(dot(dir, dir));
        const V3 nd{dir.x * dinv, dir.y * dinv, dir.z * dinv};
        const V3 oc = sub(pt, center);
        const float a = dot(nd, nd);
        const float bq = 2.0f * dot(oc, nd);
        const float c = dot(oc, oc) - 1.0f;
        const float disc = fmaf(bq, bq, -((4.0f * a) * c));
        float wgt = 0.f;
        if (disc > 0.f) {
            const float sq = sqrtf(disc);
            const float t1 = (-bq - sq) / (2.0f * a);
            const float t2 = (-bq + sq) / (2.0f * a);
            const V3 i1{fmaf(t1, nd.x, pt.x), fmaf(t1, nd.y, pt.y), fmaf(t1, nd.z, pt.z)};
            const V3 i2{fmaf(t2, nd.x, pt.x), fmaf(t2, nd.y, pt.y), fmaf(t2, nd.z, pt.z)};
            const V3 df = sub(i1, i2);
            wgt = sqrtf(dot(df, df));
        }
        for (int b = 0; b < nb; ++b) st.energy[b] = st.energy[b] * wgt;
        /* :124-170  deposit */
        const float elapsed = st.distance / 343.0f;
        const int bin = (int)roundf(elapsed * (float)P.sample_rate);
        const float cross_gain = 1.0f - P.hrtf_absorption_rate;
        const int primary = (m == -1) ? 0 : 1;
        dep.idx[0] = bin; dep.ear[0] = primary;          /* recorded even when bin >= ir_length (not deposited) */
        for (int b = 0; b < nb; ++b) dep.val[0][b] = st.energy[b];
        if (bin < P.ir_length) {
            dep.n = 1;
            if (!P.is_mono) {
                dep.n = 2;
                dep.ear[1] = 1 - primary;
                dep.idx[1] = (bin + delay < P.ir_length) ? bin + delay : bin;
                for (int b = 0; b < nb; ++b) dep.val[1][b] = st.energy[b] * cross_gain;
            }
        }
        st.depth = -1;                                   /* :147,169 */
        return;
    }
    /* :171-176  wall */
    bool diffuse = false;
    uint32_t r[4];
    if (sc > 0.f) {
        philox(P.seed, ray, (uint32_t)st.depth, 1u, r);
        diffuse = (float)(r[0] >> 8) * 0x1p-24f < sc;
    }
    if (diffuse) {
        st.dir = lambert_direction(r, dir, ng);
    } else {
        const float k = 2.0f * dot(dir, ng);
        st.dir = {fmaf(-k, ng.x, dir.x), fmaf(-k, ng.y, dir.y), fmaf(-k, ng.z, dir.z)};
    }
    for (int b = 0; b < nb; ++b) st.energy[b] = st.energy[b] * (1.0f - absorption[b]);
    st.depth++;
    /* :179  prev_position = P + 1e-3 * direction */
    st.prev = {fmaf(1e-3f, st.dir.x, pt.x), fmaf(1e-3f, st.dir.y, pt.y), fmaf(1e-3f, st.dir.z, pt.z)};
}

/* One thread of __raygen__renderFrame + its closest-hit/miss programs. */
void trace_ray(const oracle_params& P, const Scene& S, uint64_t ray, float energy0, float dist_thr, int delay,
               double* hist, RayOut& out)
{
    const int nb = P.bands;
    PathState st;
    for (int b = 0; b < nb; ++b) st.energy[b] = energy0;        /* devicePrograms.cu:208 */
    st.distance = 0.f;                                           /* :209 */
    st.prev = {P.emitter[0], P.emitter[1], P.emitter[2]};        /* :210 */
    st.depth = 0;                                                /* :211 */
    const V3 center{P.sphere_center[0], P.sphere_center[1], P.sphere_center[2]};
    st.dir = emit_direction(P.seed, ray);                        /* :216-224 */
    out.bin = -1; out.ear = 0; out.nseg = 0;
    for (int b = 0; b < nb; ++b) out.energy[b] = 0.f;
    if (!(st.dir.x != 0.f || st.dir.y != 0.f || st.dir.z != 0.f)) return; /* :230 */

    for (;;) {
        float emax = st.energy[0];
        for (int b = 1; b < nb; ++b) emax = std::max(emax, st.energy[b]);
        if (!(st.distance < dist_thr && emax > P.energy_thres && st.depth >= 0 && (uint32_t)st.depth < P.max_bounces))
            break;                                               /* :233-236 */
        out.nseg++;
        const Hit h = S.use_bvh ? closest_bvh(S.bvh, S.tris, st.prev, st.dir) : closest_brute(S.tris, st.prev, st.dir);
        if (h.id < 0) { st.depth = -1; break; }                  /* __miss__radiance :186-190 */
        const int m = S.mat[h.id];
        Deposits dep;
        shade_hit(P, nb, S.tris[h.id].p1, S.p2[h.id], S.p3[h.id], m, m >= 0 ? S.absorption + (size_t)m * nb : nullptr,
                  (m >= 0 && S.scattering) ? S.scattering[m] : 0.f, ray, center, h.u, h.v, delay, st, dep);
        if (m < 0) {
            out.bin = dep.idx[0]; out.ear = (m == -1) ? 1 : 2;
            for (int b = 0; b < nb; ++b) out.energy[b] = dep.val[0][b];
            if (hist && out.bin >= 0)
                for (int d = 0; d < dep.n; ++d)
                    for (int b = 0; b < nb; ++b)
                        hist[((size_t)dep.ear[d] * nb + b) * P.ir_length + dep.idx[d]] += (double)dep.val[d][b];
            break;
        }
    }
}

} // namespace

extern "C" {

/* Scene handle: triangles + the oracle's BVH, built once and traced many times (used
 * by bench.py's CPU baseline so that the BVH build stays outside the timed region). */
struct oracle_scene_t {
    Scene S;
    std::vector<int32_t> mat;
    std::vector<float> absorption, scattering;
    int bands;
};

void* oracle_scene_create(const float* tri_verts, const int32_t* tri_mat, int64_t n_tris, const float* absorption,
                          const float* scattering, int32_t n_mats, int32_t bands, int32_t use_bvh)
{
    if (n_tris < 0 || bands < 1 || bands > MAX_BANDS) return nullptr;
    auto* h = new oracle_scene_t;
    h->bands = bands;
    h->mat.assign(tri_mat, tri_mat + n_tris);
    h->absorption.assign(absorption, absorption + (size_t)n_mats * bands);
    if (scattering) h->scattering.assign(scattering, scattering + n_mats);
    Scene& S = h->S;
    S.tris = make_tris(tri_verts, n_tris);
    S.p2.resize(n_tris); S.p3.resize(n_tris);
    for (int64_t i = 0; i < n_tris; ++i) {
        const float* q = tri_verts + 9 * i;
        S.p2[i] = {q[3], q[4], q[5]}; S.p3[i] = {q[6], q[7], q[8]};
    }
    S.mat = h->mat.data(); S.absorption = h->absorption.data();
    S.scattering = h->scattering.empty() ? nullptr : h->scattering.data();
    S.use_bvh = use_bvh != 0;
    if (S.use_bvh) S.bvh = build_bvh(S.tris);
    return h;
}

void oracle_scene_destroy(void* h) { delete (oracle_scene_t*)h; }

int64_t oracle_trace_scene(void* handle, const oracle_params* p, int64_t ray_begin, int64_t n_rays, int32_t n_threads,
                           double* hist, int32_t* rec_bin, int32_t* rec_ear, float* rec_energy, int32_t* rec_nseg)
{
    if (!handle || !p || n_rays < 0) return -1;
    const oracle_scene_t* h = (const oracle_scene_t*)handle;
    if (p->bands != h->bands) return -1;
    const oracle_params P = *p;
    const Scene& S = h->S;

    /* devicePrograms.cu:208  base_power / ((x*y*z) * 4.18879020478) in double, narrowed */
    const int n_total = P.size_x * P.size_y * P.size_z;
    const float energy0 = (float)((double)P.base_power / ((double)n_total * 4.18879020478));
    /* :227-228 */
    const int ir_sec = std::max(1, std::min(P.ir_length / P.sample_rate, 999));
    const float dist_thr = (float)(ir_sec * 343 + 1);
    /* :125 */
    const int delay = (int)((double)P.sample_rate * 0.00044);

    const size_t hsz = (size_t)2 * P.bands * P.ir_length;
    if (hist) std::memset(hist, 0, hsz * sizeof(double));
    const int nt = std::max(1, n_threads);
    std::vector<std::vector<double>> priv(nt > 1 && hist ? nt : 0);
    std::vector<int64_t> segs(nt, 0);
    std::atomic<int64_t> next{0};
    const int64_t chunk = 4096;
    auto work = [&](int tid) {
        double* hh = hist;
        if (nt > 1 && hist) { priv[tid].assign(hsz, 0.0); hh = priv[tid].data(); }
        for (;;) {
            const int64_t b = next.fetch_add(chunk);
            if (b >= n_rays) break;
            const int64_t e = std::min(n_rays, b + chunk);
            for (int64_t i = b; i < e; ++i) {
                RayOut o;
                trace_ray(P, S, (uint64_t)(ray_begin + i), energy0, dist_thr, delay, hh, o);
                segs[tid] += o.nseg;
                if (rec_bin) rec_bin[i] = o.bin;
                if (rec_ear) rec_ear[i] = o.ear;
                if (rec_nseg) rec_nseg[i] = o.nseg;
                if (rec_energy) for (int k = 0; k < P.bands; ++k) rec_energy[i * P.bands + k] = o.energy[k];
            }
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto& t : th) t.join();
        if (hist) for (int t = 0; t < nt; ++t) for (size_t i = 0; i < hsz; ++i) hist[i] += priv[t][i];
    }
    int64_t total = 0;
    for (auto s : segs) total += s;
    return total;
}

int64_t oracle_trace(const oracle_params* p, const float* tri_verts, const int32_t* tri_mat, int64_t n_tris,
                     const float* absorption, const float* scattering, int32_t n_mats, int64_t ray_begin,
                     int64_t n_rays, int32_t use_bvh, int32_t n_threads, double* hist, int32_t* rec_bin,
                     int32_t* rec_ear, float* rec_energy, int32_t* rec_nseg)
{
    if (!p) return -1;
    void* h = oracle_scene_create(tri_verts, tri_mat, n_tris, absorption, scattering, n_mats, p->bands, use_bvh);
    if (!h) return -1;
    const int64_t r = oracle_trace_scene(h, p, ray_begin, n_rays, n_threads, hist, rec_bin, rec_ear, rec_energy, rec_nseg);
    oracle_scene_destroy(h);
    return r;
}

/* One closest-hit invocation (single band) in the flat form oracle/ref_device_shim.cpp's ref_closesthit uses, for
 * the pin test against the reference's own __closesthit__radiance: prd8 = {remaining_factor, distance,
 * prev_position[3], direction[3]} in/out. */
void oracle_shade_hit(const float* tri9, float mat_absorption, const float* ray_dir, float u, float v,
                      const float* sphere_center, int32_t sample_rate, float hrtf, int32_t is_mono, int32_t ir_length,
                      float* prd8, int32_t* depth, int32_t* n_dep, int32_t* dep_ear, int32_t* dep_idx, float* dep_val)
{
    oracle_params P{};
    P.sample_rate = sample_rate; P.hrtf_absorption_rate = hrtf; P.is_mono = is_mono; P.ir_length = ir_length; P.bands = 1;
    PathState st;
    st.energy[0] = prd8[0]; st.distance = prd8[1];
    st.prev = {prd8[2], prd8[3], prd8[4]};
    st.dir = {ray_dir[0], ray_dir[1], ray_dir[2]};      /* the traced ray's direction is prd.direction (:238,:245) */
    (void)prd8[5];
    st.depth = *depth;
    const int m = mat_absorption == -1.f ? -1 : (mat_absorption == -2.f ? -2 : 0);
    const int delay = (int)((double)sample_rate * 0.00044);
    Deposits dep;
    shade_hit(P, 1, {tri9[0], tri9[1], tri9[2]}, {tri9[3], tri9[4], tri9[5]}, {tri9[6], tri9[7], tri9[8]}, m, &mat_absorption, 0.f,
              0, {sphere_center[0], sphere_center[1], sphere_center[2]}, u, v, delay, st, dep);
    prd8[0] = st.energy[0]; prd8[1] = st.distance;
    prd8[2] = st.prev.x; prd8[3] = st.prev.y; prd8[4] = st.prev.z;
    prd8[5] = st.dir.x; prd8[6] = st.dir.y; prd8[7] = st.dir.z;
    *depth = st.depth;
    *n_dep = (m < 0 && dep.idx[0] >= 0) ? dep.n : 0;
    for (int i = 0; i < 2; ++i) { dep_ear[i] = dep.ear[i]; dep_idx[i] = dep.idx[i]; dep_val[i] = dep.val[i][0]; }
}

void oracle_ray_direction(uint64_t seed, uint64_t ray_id, float* dir3)
{
    const V3 d = emit_direction(seed, ray_id);
    dir3[0] = d.x; dir3[1] = d.y; dir3[2] = d.z;
}

void oracle_philox(uint64_t seed, uint64_t ray_id, uint32_t bounce, uint32_t purpose, uint32_t* out4)
{
    philox(seed, ray_id, bounce, purpose, out4);
}

int64_t oracle_closest_hit(const float* tri_verts, int64_t n_tris, const float* org3, const float* dir3, float* t,
                           float* u, float* v)
{
    const std::vector<Tri> tris = make_tris(tri_verts, n_tris);
    const Hit h = closest_brute(tris, {org3[0], org3[1], org3[2]}, {dir3[0], dir3[1], dir3[2]});
    if (h.id >= 0) { *t = h.t; *u = h.u; *v = h.v; }
    return h.id;
}

void oracle_finalize_ir(const double* hist, int32_t bands, int32_t ir_length, int32_t is_mono, float* ir_left,
                        float* ir_right)
{
    const size_t n = (size_t)bands * ir_length;
    for (size_t i = 0; i < n; ++i) {
        const float l = (float)hist[i], r = (float)hist[n + i];
        if (is_mono) { const float s = l + r; ir_left[i] = s; ir_right[i] = s; }
        else { ir_left[i] = l; ir_right[i] = r; }
    }
}

} // extern "C"
