// Build of the far-field direction grid (fargrid.cuh) at scene upload: per-triangle far-field constants,
// then the strips of all triangles rasterised into the 6 K^2 direction cells (count, scan, fill).
// Replaces, for rays that leave the scene, the scan over every triangle's filter record of round 1;
// the test it prepares is the reference's own (Raytracer.cpp:348-409), see fargrid.cuh.
#include "fargrid.cuh"
#include "build.h"
#include <cstdio>
#include <cmath>

namespace rt580 {

struct FgParams {
    int K;
    float h;            // cell edge on the face plane: 2 / K
    float r_c;          // angular radius of a cell (upper bound, + the rounding of the lookup)
    float diag;         // bound on |O - v| for in-scene origins O and scene points v
    float extent;
    float ob_lo[3], ob_hi[3], cam[3];
};

// per-triangle constants the rasteriser needs, recomputed from the primitive record
struct FgTri {
    V3 N, k1, k2, m;
    float mlen, elen, Ti, dmax, eps_o, thr1;
    bool ok;
};

__device__ __forceinline__ FgTri fg_tri(const PrimRec* __restrict__ prims, const float2* __restrict__ fgB, int i, const FgParams& fp)
{
    FgTri t; t.ok = false;
    const float2 B = fgB[i];
    if (!(B.x > 0.f) || B.x < FG_WIDE_FACTOR * fp.diag) return t;            // no far field / wide list
    const float4 ra = prims[i].a, rb = prims[i].b, rc = prims[i].c, rd = prims[i].d;
    const V3 v0 = mk(ra.x, ra.y, ra.z), v1 = mk(rb.x, rb.y, rb.z), v2 = mk(rc.x, rc.y, rc.z);
    t.N = mk(rd.x, rd.y, rd.z);
    const V3 E1 = v1 - v0, E2 = v2 - v0, e = v2 - v1;
    // in-plane inward normals of the wedge at v0: gamma >= 0 <=> p.(N x E1) >= 0, beta >= 0 <=> p.(E2 x N) >= 0  (cpp:393-394)
    t.k1 = normalize(cross(t.N, E1));
    t.k2 = normalize(cross(E2, t.N));
    t.m = cross(e, t.N);
    t.mlen = length(t.m); t.elen = length(e);
    t.Ti = B.x; t.dmax = B.y;
    t.eps_o = 1.02f * fp.diag / (t.Ti - fp.diag);
    t.thr1 = t.dmax / t.Ti * 1.00001f + FG_ND_SLACK;
    t.ok = t.mlen > 0.f && t.elen > 0.f;
    return t;
}

// Lower bound of the ray parameter of a far-field hit of triangle `t` by a ray whose direction lies in the cell with
// centre direction c (unit) and angular radius r_c; 0: the cell is outside the wedge.  (fargrid.cuh header.)
__device__ __forceinline__ float fg_cell_T(const FgTri& t, V3 c, const FgParams& fp)
{
    const float eps = t.eps_o + fp.r_c + 2e-6f;
    if (dot(c, t.k1) < -(eps + 2e-6f) || dot(c, t.k2) < -(eps + 2e-6f)) return 0.f;
    const float g = fabsf(c.y * c.z * t.N.x) + fabsf(c.z * c.x * t.N.y) + fabsf(c.x * c.y * t.N.z);
    const float num = fabsf(dot(c, t.m)) - eps * t.mlen - (float)(9.3 * FG_U) * t.elen;
    float T = t.Ti;
    if (num > 0.f) {
        const float den = (float)(6.0001 * FG_U) * (g + 3.51f * eps);
        const float TL = num / den * 0.98f;
        T = fmaxf(T, (TL - fp.diag) * 0.999999f);
    }
    return T;
}

// One warp per triangle: the rows (or columns) of the six faces are dealt to the lanes.
template <bool FILL>
__global__ void __launch_bounds__(256)
k_fg_raster(const PrimRec* __restrict__ prims, const float2* __restrict__ fgB, int n_all, FgParams fp,
            unsigned int* __restrict__ counts, const unsigned long long* __restrict__ start, uint32_t* __restrict__ entries,
            unsigned int* __restrict__ cell_tmin)
{
    const int i = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n_all) return;
    const FgTri t = fg_tri(prims, fgB, i, fp);
    if (!t.ok) return;
    const int K = fp.K;
    const float Wd = (t.thr1 + fp.r_c) * 1.7338f;               // band on the face plane: |N.p| <= (thr + r_c) |p|, |p| <= sqrt(3)
    for (int face = 0; face < 6; face++) {
        const int a = face >> 1, ua = (a + 1) % 3, va = (a + 2) % 3;
        const float sg = (face & 1) ? -1.f : 1.f;
        const float Nu = fg_axis(t.N, ua), Nv = fg_axis(t.N, va), Nw = fg_axis(t.N, a) * sg;
        if (fabsf(Nw) > fabsf(Nu) + fabsf(Nv) + Wd) continue;   // the great circle misses this face
        const bool rows = fabsf(Nu) >= fabsf(Nv);               // iterate rows (v), solve for u; else the other way round
        const float Ns = rows ? Nu : Nv, No = rows ? Nv : Nu;   // solved / iterated coefficient
        const float inv = 1.0f / Ns;
        const float halfw = Wd * fabsf(inv);
        // Rows worth visiting: with o the iterated coordinate, the strip's centre line is s(o) = -(No o + Nw) / Ns; it must lie on
        // the face (|s| <= 1 + halfw) and inside the widened wedge (p.k >= -m |p| for both in-plane normals, p = (o, s(o), 1) in
        // face coordinates): four affine inequalities c0 + c1 o >= 0, intersected with [-1, 1].  A superset: the cells are
        // tested one by one below.
        float o_lo = -1.f, o_hi = 1.f;
        {
            const float s0 = -Nw * inv, s1 = -No * inv;                       // s(o) = s0 + s1 o
            auto clip = [&](float c0, float c1) {                             // keep c0 + c1 o >= 0
                if (c1 > 0.f) o_lo = fmaxf(o_lo, -c0 / c1);
                else if (c1 < 0.f) o_hi = fminf(o_hi, -c0 / c1);
                else if (c0 < 0.f) o_hi = -2.f;
            };
            const float slack = 1.f + halfw + 2.f * fp.h;
            clip(slack - s0, -s1); clip(slack + s0, s1);                      // |s(o)| <= 1 + halfw
            const float m = (t.eps_o + 3.f * fp.r_c + t.thr1 + 1e-5f) * 1.7338f + halfw;
            const V3 kk[2] = { t.k1, t.k2 };
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const float ku = fg_axis(kk[q], ua), kv = fg_axis(kk[q], va), kw = fg_axis(kk[q], a) * sg;
                const float ks = rows ? ku : kv, ko = rows ? kv : ku;        // coefficient of the solved / iterated coordinate
                clip(m + kw + ks * s0, ko + ks * s1);
            }
        }
        if (o_hi < o_lo) continue;
        const int r_lo = max(0, (int)floorf((o_lo + 1.f) / fp.h) - 1), r_hi = min(K - 1, (int)floorf((o_hi + 1.f) / fp.h) + 1);
        for (int r = r_lo + lane; r <= r_hi; r += 32) {
            const float oc = ((float)r + 0.5f) * fp.h - 1.f;    // centre of the row in the iterated coordinate
            const float line = -(No * oc + Nw) * inv;
            const float lo = line - halfw, hi = line + halfw;
            if (hi < -1.f || lo > 1.f) continue;
            const int s0 = max(0, (int)floorf((lo + 1.f) / fp.h - 0.5f)), s1 = min(K - 1, (int)floorf((hi + 1.f) / fp.h + 0.5f));
            for (int s = s0; s <= s1; s++) {
                const float sc_ = ((float)s + 0.5f) * fp.h - 1.f;
                const float pu = rows ? sc_ : oc, pv = rows ? oc : sc_;
                const float plen = sqrtf(pu * pu + pv * pv + 1.f);
                const float Np = Nu * pu + Nv * pv + Nw;
                if (fabsf(Np) > (t.thr1 + fp.r_c) * plen * 1.0001f) continue;      // outside the widest band this triangle has
                V3 c;
                { float cc[3]; cc[ua] = pu / plen; cc[va] = pv / plen; cc[a] = sg / plen; c = mk(cc[0], cc[1], cc[2]); }
                const float T = fg_cell_T(t, c, fp);
                if (!(T > 0.f)) continue;
                const float thr = t.dmax / T * 1.00001f + FG_ND_SLACK;
                if (fabsf(Np) > (thr + fp.r_c) * plen * 1.0001f) continue;
                const int iu = rows ? s : r, iv = rows ? r : s;
                const size_t cell = ((size_t)face * K + iv) * K + iu;
                if (!FILL) atomicAdd(counts + cell, 1u);
                else {
                    int k6 = (int)floorf(4.0f * log2f(T / t.Ti) - 1e-3f);
                    k6 = max(0, min(63, k6));
                    const unsigned pos = atomicAdd(counts + cell, 1u);
                    entries[start[cell] + pos] = ((unsigned)k6 << FG_ID_BITS) | (unsigned)i;
                    atomicMin(cell_tmin + cell, __float_as_uint(fg_entry_T(t.Ti, (unsigned)k6)));    // (positive floats order like their bits)
                }
            }
        }
    }
}

// Per primitive: (N, D) and the direction-independent bound T of the ray parameter of a far-field hit, from the filter
// record of the build (bvh_build.cu: thr_old = 4E / T_old, rigorous, round 1) tightened by the bound of fargrid.cuh.
__global__ void __launch_bounds__(256)
k_fg_setup(const PrimRec* __restrict__ prims, const float4* __restrict__ far_old, int n_all, FgParams fp,
           float4* __restrict__ fgA, float2* __restrict__ fgB, uint32_t* __restrict__ wide, uint32_t* __restrict__ sph,
           unsigned int* __restrict__ counters)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_all) return;
    const float4 ra = prims[i].a, rb = prims[i].b, rc = prims[i].c, rd = prims[i].d;
    const float thr_old = far_old[i].w;
    const bool tri = !(__float_as_int(rd.w) & RT_PRIM_SPHERE);
    fgA[i] = tri ? make_float4(rd.x, rd.y, rd.z, ra.w) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (!tri) sph[atomicAdd(counters + 2, 1u)] = (unsigned)i;
    float T = -1.f, dmax = 0.f;
    if (tri && thr_old > 0.f && thr_old < 4.0f) {
        const double E = fp.extent, diag = fp.diag, u = FG_U;
        const double T_old = 4.0 * E / (double)thr_old * (1.0 - 1e-6);
        const double v0[3] = { ra.x, ra.y, ra.z }, v1[3] = { rb.x, rb.y, rb.z }, v2[3] = { rc.x, rc.y, rc.z }, N[3] = { rd.x, rd.y, rd.z };
        // |N.O + D| over in-scene origins: the corners of the origin box and the camera
        double dm = fabs(N[0] * (fp.cam[0] - v0[0]) + N[1] * (fp.cam[1] - v0[1]) + N[2] * (fp.cam[2] - v0[2]));
        for (int c = 0; c < 8; c++) {
            const double x = (c & 1) ? fp.ob_hi[0] : fp.ob_lo[0], y = (c & 2) ? fp.ob_hi[1] : fp.ob_lo[1], z = (c & 4) ? fp.ob_hi[2] : fp.ob_lo[2];
            dm = fmax(dm, fabs(N[0] * (x - v0[0]) + N[1] * (y - v0[1]) + N[2] * (z - v0[2])));
        }
        // + the float evaluation of -(dot(N, O) + D) with D = -dot(N, v0) (cpp:377, 381): <= 8u (|N.O| + |N.v0|) terms
        dmax = (float)(dm * 1.00001 + 16.0 * u * 3.5 * E + 1e-30);
        double Tn = T_old;
        bool degenerate = true;
        {
            double e[3], a1[3], a2[3], E1[3], E2[3];
            for (int k = 0; k < 3; k++) { e[k] = v2[k] - v1[k]; a1[k] = v0[k] - v1[k]; a2[k] = v0[k] - v2[k]; E1[k] = -a1[k]; E2[k] = -a2[k]; }
            auto crs = [](const double* a, const double* b, double* o) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; };
            auto len = [](const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); };
            double m[3], x1[3], x2[3], k1[3], k2[3];
            crs(e, N, m); crs(a1, e, x1); crs(a2, e, x2); crs(N, E1, k1); crs(E2, N, k2);
            const double el = len(e), l1 = len(a1), l2 = len(a2);
            // the wedge test of the rasteriser takes N perpendicular to the two edges at v0 up to 2 %: slivers whose float
            // normal is off by more stay out of the direction index (wide list)
            degenerate = !(el > 0 && l1 > 0 && l2 > 0 && len(k1) > 0.98 * l1 && len(k2) > 0.98 * l2 && len(m) > 0.98 * el);
            if (!degenerate && T_old > 2.2 * diag) {
                // the direction of P seen from v1 and the ray's own direction differ by at most mu (|O - v1| <= diag, |P - v1| >= T_old)
                const double mu = asin(fmin(diag / (T_old - diag), 1.0)) * 1.002 + 2e-6;
                auto dot3 = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
                const double me[3] = { -e[0], -e[1], -e[2] };
                const double ang1 = atan2(len(x1), dot3(a1, e)), ang2 = atan2(len(x2), dot3(a2, me));   // angles at v1, v2
                // over the widened wedge the angle between w and e runs from ang2 - mu to pi - ang1 + mu: sin is concave there
                if (ang1 > mu && ang2 > mu) {
                    const double smin = fmin(sin(ang1 - mu), sin(ang2 - mu));
                    const double TL = (len(m) * smin * (1.0 - 1e-4) - 9.3 * u * el) / (6.0001 * u * 0.57741 * 1.000001) * 0.98;
                    Tn = fmax(Tn, (TL - diag) * (1.0 - 1e-6));
                }
            }
        }
        T = (float)(Tn * (1.0 - 1e-6));
        if (degenerate) T = fminf(T, 0.99f * FG_WIDE_FACTOR * fp.diag);
        if (T < FG_WIDE_FACTOR * fp.diag) {
            const unsigned slot = atomicAdd(counters, 1u);
            wide[slot] = (unsigned)i;
        }
    }
    fgB[i] = make_float2(T, dmax);
    if (T > 0.f) atomicMin(counters + 1, __float_as_uint(T));
}

// ---- exclusive scan uint32 -> uint64 (cells of the grid; totals beyond 2^32 for 10^7 triangles) ----
#define FGS_ITEMS 8
__global__ void __launch_bounds__(256)
k_fg_scan_reduce(const unsigned int* __restrict__ in, size_t n, unsigned long long* __restrict__ bsum)
{
    __shared__ unsigned long long ws[8];
    const size_t base = ((size_t)blockIdx.x * 256 + threadIdx.x) * FGS_ITEMS;
    unsigned long long s = 0;
    for (int k = 0; k < FGS_ITEMS; k++) if (base + k < n) s += in[base + k];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned long long t = 0; for (int k = 0; k < 8; k++) t += ws[k]; bsum[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024)
k_fg_scan_top(unsigned long long* __restrict__ bsum, unsigned n_blocks, unsigned long long* __restrict__ total)
{
    __shared__ unsigned long long ws[32];
    __shared__ unsigned long long carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0ull;
    __syncthreads();
    for (unsigned b0 = 0; b0 < n_blocks; b0 += 1024) {
        const unsigned b = b0 + threadIdx.x;
        const unsigned long long v = b < n_blocks ? bsum[b] : 0ull;
        unsigned long long incl = v;
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) ws[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = ws[lane];
            for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
            ws[lane] = w;
        }
        __syncthreads();
        const unsigned long long carry = carry_s;
        if (b < n_blocks) bsum[b] = carry + incl - v + (warp ? ws[warp - 1] : 0ull);
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + ws[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}
__global__ void __launch_bounds__(256)
k_fg_scan_apply(const unsigned int* __restrict__ in, size_t n, const unsigned long long* __restrict__ boff,
                unsigned long long* __restrict__ out)
{
    __shared__ unsigned long long ws[8];
    const size_t base = ((size_t)blockIdx.x * 256 + threadIdx.x) * FGS_ITEMS;
    unsigned int v[FGS_ITEMS]; unsigned long long s = 0;
    for (int k = 0; k < FGS_ITEMS; k++) { v[k] = (base + k < n) ? in[base + k] : 0u; s += v[k]; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long incl = s;
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    unsigned long long woff = 0;
    for (int k = 0; k < warp; k++) woff += ws[k];
    unsigned long long excl = boff[blockIdx.x] + woff + incl - s;
    for (int k = 0; k < FGS_ITEMS; k++) { if (base + k < n) out[base + k] = excl; excl += v[k]; }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { snprintf(err, errlen, "%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return false; } } while (0)

int fg_default_K(long long n_tris)
{
    if (n_tris < 2048) return 0;                  // the O(n) scan is cheap there
    if (n_tris < (1 << 14)) return 64;
    if (n_tris < (1 << 16)) return 128;
    if (n_tris < (1 << 18)) return 256;
    if (n_tris < 600000) return 512;
    if (n_tris <= 2500000) return 1024;
    return 512;                                   // memory: entries ~ 1.5 K n
}

bool fg_build(const FgBuildInput& in, FgBuildOutput* out, cudaStream_t stream, char* err, size_t errlen)
{
    out->K = 0; out->n_wide = 0; out->n_sph = 0; out->n_entries = 0; out->t_min = 3.0e38f;
    const int n = in.n_all;
    if (n <= 0) return true;
    FgParams fp{};
    fp.K = in.K; fp.h = in.K > 0 ? 2.0f / (float)in.K : 0.f;
    fp.r_c = 0.7072f * fp.h * 1.001f + 1e-6f;
    fp.extent = in.extent;
    double d2 = 0;
    for (int k = 0; k < 3; k++) {
        fp.ob_lo[k] = in.ob_lo[k]; fp.ob_hi[k] = in.ob_hi[k]; fp.cam[k] = in.cam[k];
        const double lo = fmin((double)in.ob_lo[k], (double)in.cam[k]), hi = fmax((double)in.ob_hi[k], (double)in.cam[k]);
        d2 += (hi - lo) * (hi - lo);
    }
    fp.diag = (float)(sqrt(d2) * 1.00001 + 1e-3);
    out->diag = fp.diag;
    CK(cudaMemsetAsync(in.counters, 0, 4 * sizeof(unsigned int), stream));
    CK(cudaMemsetAsync(in.counters + 1, 0x7f, sizeof(unsigned int), stream));
    k_fg_setup<<<(n + 255) / 256, 256, 0, stream>>>(in.prims, in.far_old, n, fp, in.fgA, in.fgB, in.wide, in.sph, in.counters);
    unsigned int hc[4];
    CK(cudaMemcpyAsync(hc, in.counters, sizeof hc, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    out->n_wide = (int)hc[0]; out->n_sph = (int)hc[2];
    { float t; memcpy(&t, &hc[1], 4); out->t_min = (hc[1] == 0x7f7f7f7fu) ? 3.0e38f : t; }
    if (in.K <= 0) return true;
    if (n > FG_MAX_PRIMS) { snprintf(err, errlen, "far-field grid: %d primitives exceed the %d-bit entry index", n, FG_ID_BITS); return false; }
    const size_t n_cells = (size_t)6 * in.K * in.K;
    CK(in.counts->ensure(n_cells + 1, 0, stream));
    CK(in.start->ensure(n_cells + 1, 0, stream));
    const unsigned n_blocks = (unsigned)((n_cells + 256 * FGS_ITEMS - 1) / (256 * FGS_ITEMS));
    CK(in.bsum->ensure((size_t)n_blocks + 2, 0, stream));
    CK(cudaMemsetAsync(in.counts->p, 0, sizeof(unsigned int) * (n_cells + 1), stream));
    const unsigned grid = (unsigned)(((size_t)n * 32 + 255) / 256);
    k_fg_raster<false><<<grid, 256, 0, stream>>>(in.prims, in.fgB, n, fp, in.counts->p, nullptr, nullptr, nullptr);
    k_fg_scan_reduce<<<n_blocks, 256, 0, stream>>>(in.counts->p, n_cells, in.bsum->p);
    k_fg_scan_top<<<1, 1024, 0, stream>>>(in.bsum->p, n_blocks, in.start->p + n_cells);
    k_fg_scan_apply<<<n_blocks, 256, 0, stream>>>(in.counts->p, n_cells, in.bsum->p, in.start->p);
    unsigned long long total = 0;
    CK(cudaMemcpyAsync(&total, in.start->p + n_cells, sizeof total, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    CK(cudaGetLastError());
    CK(in.entries->ensure((size_t)total + 1, 0, stream));
    CK(cudaMemsetAsync(in.counts->p, 0, sizeof(unsigned int) * (n_cells + 1), stream));
    CK(in.cell_tmin->ensure(n_cells + 1, 0, stream));
    CK(cudaMemsetAsync(in.cell_tmin->p, 0xff, sizeof(unsigned int) * (n_cells + 1), stream));
    k_fg_raster<true><<<grid, 256, 0, stream>>>(in.prims, in.fgB, n, fp, in.counts->p, in.start->p, in.entries->p, in.cell_tmin->p);
    CK(cudaStreamSynchronize(stream));
    CK(cudaGetLastError());
    out->K = in.K; out->n_entries = total;
    return true;
}

}  // namespace rt580
