// raster.cu — vb200_render_depth_batch: a compute-shader-free CUDA z-buffer rasteriser for render_depth.
// Replaces feh::Renderer::{SetCamera, SetMesh, RenderDepth} (render/renderer.cpp:232-351) and its GLSL
// vertex shader (render/shaders/basic_mvp.vert:10): the OpenGL context, FBO and synchronous glReadPixels
// per map become three launches for a whole batch of meshes:
//   k_clear   z-buffer <- 2^24-1 (glClear depth 1)
//   k_tris    setup by one thread per (mesh, triangle): float vertex stage, near/far clipping, sub-pixel
//             snapping; then the WARP rasterises its 32 triangles together: the pixels of all their boxes form
//             one flat work list the 32 lanes stride over (integer edge functions, 24-bit z with atomicMin =
//             GL_LESS), so a warp is busy whatever the mix of triangle sizes.  Clipped polygons and triangles
//             with a very large pixel box are queued instead
//   k_big     one warp per queued triangle
//   k_resolve uint32 z -> float depth (what glReadPixels(GL_DEPTH_COMPONENT, GL_FLOAT) hands back)
// The arithmetic is written with explicitly rounded intrinsics (no FMA contraction) and must match the
// canonical rules restated in oracle/raster_oracle.c bit for bit.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace vb {

namespace {

constexpr unsigned kZMax = 16777215u;
constexpr int kSub = 256;
constexpr double kClamp = 536870912.0;  // 2^29 sub-pixels
constexpr int kFlatMax = 2048;          // largest pixel box rasterised inside its warp's flat work list
constexpr int kCoordSafe = 1 << 22;     // |X|, |Y| below this (and H, W <= 8192): edge-function operands fit 32 bits

struct MeshDesc {
    float mvp[16];  // column-major (P*V)*M
    int64_t v_begin, f_begin;
    int nv, nf;
};

struct TriSetup {
    int X0, Y0, X1, Y1, X2, Y2;
    double z0, z1, z2;
    int mesh;
    int pad;
};

struct ClipV { double x, y, z, w; };

__device__ __forceinline__ int clip_poly(const ClipV *in, int n, ClipV *out, int plane) {
    int m = 0;
    for (int i = 0; i < n; i++) {
        const ClipV a = in[i], b = in[(i + 1) % n];
        double da = plane == 0 ? __dadd_rn(a.w, a.z) : __dsub_rn(a.w, a.z);
        double db = plane == 0 ? __dadd_rn(b.w, b.z) : __dsub_rn(b.w, b.z);
        bool ia = da >= 0.0, ib = db >= 0.0;
        if (ia) out[m++] = a;
        if (ia != ib) {
            double t = __ddiv_rn(da, __dsub_rn(da, db));
            ClipV c;
            c.x = __dadd_rn(a.x, __dmul_rn(t, __dsub_rn(b.x, a.x)));
            c.y = __dadd_rn(a.y, __dmul_rn(t, __dsub_rn(b.y, a.y)));
            c.z = __dadd_rn(a.z, __dmul_rn(t, __dsub_rn(b.z, a.z)));
            c.w = __dadd_rn(a.w, __dmul_rn(t, __dsub_rn(b.w, a.w)));
            out[m++] = c;
        }
    }
    return m;
}

__device__ __forceinline__ int snap(double v) {
    double s = __dmul_rn(v, (double)kSub);
    if (!(s > -kClamp)) s = -kClamp;
    if (s > kClamp) s = kClamp;
    return (int)__double2ll_rn(s);
}

__device__ __forceinline__ bool top_left(long long dx, long long dy) { return dy < 0 || (dy == 0 && dx > 0); }

struct TriRaster {
    long long X0, Y0, X1, Y1, X2, Y2;
    long long dx0, dy0, dx1, dy1, dx2, dy2;
    double z0, z1, z2, a2;
    bool tl0, tl1, tl2;
    int i0, i1, j0, j1;
};

// returns false for degenerate / off-screen triangles
__device__ __forceinline__ bool tri_prepare(const TriSetup &s, int H, int W, TriRaster &t) {
    long long X0 = s.X0, Y0 = s.Y0, X1 = s.X1, Y1 = s.Y1, X2 = s.X2, Y2 = s.Y2;
    double z0 = s.z0, z1 = s.z1, z2 = s.z2;
    long long area2 = (X1 - X0) * (Y2 - Y0) - (X2 - X0) * (Y1 - Y0);
    if (area2 == 0) return false;
    if (area2 < 0) {
        long long tt;
        double tz;
        tt = X1; X1 = X2; X2 = tt;
        tt = Y1; Y1 = Y2; Y2 = tt;
        tz = z1; z1 = z2; z2 = tz;
        area2 = -area2;
    }
    long long minX = min(X0, min(X1, X2)), maxX = max(X0, max(X1, X2));
    long long minY = min(Y0, min(Y1, Y2)), maxY = max(Y0, max(Y1, Y2));
    static_assert(kSub == 256, "the pixel box uses >> 8 as floor division by kSub");
    long long i0 = (minX - 128 + (kSub - 1)) >> 8, i1 = (maxX - 128) >> 8;  // arithmetic shift = floor division
    long long j0 = (minY - 128 + (kSub - 1)) >> 8, j1 = (maxY - 128) >> 8;
    i0 = max(i0, 0ll); j0 = max(j0, 0ll);
    i1 = min(i1, (long long)W - 1); j1 = min(j1, (long long)H - 1);
    if (i0 > i1 || j0 > j1) return false;
    t.X0 = X0; t.Y0 = Y0; t.X1 = X1; t.Y1 = Y1; t.X2 = X2; t.Y2 = Y2;
    t.dx0 = X2 - X1; t.dy0 = Y2 - Y1; t.dx1 = X0 - X2; t.dy1 = Y0 - Y2; t.dx2 = X1 - X0; t.dy2 = Y1 - Y0;
    t.tl0 = top_left(t.dx0, t.dy0); t.tl1 = top_left(t.dx1, t.dy1); t.tl2 = top_left(t.dx2, t.dy2);
    t.z0 = z0; t.z1 = z1; t.z2 = z2;
    t.a2 = (double)area2;
    t.i0 = (int)i0; t.i1 = (int)i1; t.j0 = (int)j0; t.j1 = (int)j1;
    return true;
}

__device__ __forceinline__ void shade_pixel(const TriRaster &t, int i, int j, long long E0, long long E1,
                                            long long E2, unsigned *zbuf, int W) {
    if (!(E0 > 0 || (E0 == 0 && t.tl0))) return;
    if (!(E1 > 0 || (E1 == 0 && t.tl1))) return;
    if (!(E2 > 0 || (E2 == 0 && t.tl2))) return;
    double num = __dadd_rn(__dadd_rn(__dmul_rn((double)E0, t.z0), __dmul_rn((double)E1, t.z1)),
                           __dmul_rn((double)E2, t.z2));
    double z = __ddiv_rn(num, t.a2);
    double qd = __dmul_rn(z, (double)kZMax);
    long long q = __double2ll_rn(qd);
    if (!(qd > 0.0)) q = 0;
    if (q > (long long)kZMax) q = kZMax;
    atomicMin(zbuf + (size_t)j * W + i, (unsigned)q);  // GL_LESS against the stored minimum
}

__device__ __forceinline__ void edge_at(const TriRaster &t, long long px, long long py, long long &E0,
                                        long long &E1, long long &E2) {
    E0 = t.dx0 * (py - t.Y1) - t.dy0 * (px - t.X1);
    E1 = t.dx1 * (py - t.Y2) - t.dy1 * (px - t.X2);
    E2 = t.dx2 * (py - t.Y0) - t.dy2 * (px - t.X0);
}

__global__ void __launch_bounds__(256) k_clear(unsigned *__restrict__ z, int64_t n) {
    // grid-stride, 16 bytes per store; the tail (and an odd caller-owned device pointer) one word at a time
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    if ((reinterpret_cast<uintptr_t>(z) & 15) != 0) {
        for (int64_t i = tid; i < n; i += nth) z[i] = kZMax;
        return;
    }
    const int64_t n4 = n >> 2;
    uint4 *z4 = reinterpret_cast<uint4 *>(z);
    for (int64_t i = tid; i < n4; i += nth) z4[i] = make_uint4(kZMax, kZMax, kZMax, kZMax);
    for (int64_t i = 4 * n4 + tid; i < n; i += nth) z[i] = kZMax;
}

// One triangle of a warp's flat work list (shared memory, 112 B): everything a lane needs to test pixel number
// q of the triangle's pixel box.  Only triangles whose snapped coordinates stay below kCoordSafe get one, so
// every edge function is C + A*col + B*row with 32-bit steps (two IMAD.WIDE) — the same int64 values
// as TriRaster's, shifted by the top-left rule: E' = E + tl - 1, so that "inside" is E' >= 0 for all three.
struct __align__(16) TriRec {
    long long C0, C1;
    long long C2;
    int A0, B0;
    int A1, B1, A2, B2;
    int i0, j0, bw;
    unsigned magic;  // ceil(2^32 / bw): q / bw == __umulhi(q, magic) for q < 2^21, bw in 2..2048 (checked exhaustively)
    int start;       // first flat index of this triangle's pixels in the warp's list
    int mesh_tl;     // mesh << 3 | top-left flags of the three edges
    double c;        // (2^24-1) / a2: the quick depth quantisation below
    double z0, z1;
    double z2, a2;
};
static_assert(sizeof(TriRec) == 112, "TriRec layout");

// A covered pixel waiting for its depth: the warp compacts them so that the depth arithmetic (a third of the
// box pixels are covered on the chair batch) runs with all lanes busy.
struct __align__(16) CovItem {
    long long E0, E1;  // biased edge functions E'
    long long E2;
    int k;             // triangle (lane) of the warp's list
    int pix;           // j * W + i
};

#ifndef VB_TRIS_TPB
#define VB_TRIS_TPB 128
#endif
constexpr int kTrisTpb = VB_TRIS_TPB;  // the warps of a block are independent (no block barrier)

// Clipped polygons (rare): clip, snap, and queue the fan triangles for k_big.
__device__ __noinline__ void tri_setup_clipped(float4 c0, float4 c1, float4 c2, int mesh, int H, int W,
                                               TriSetup *__restrict__ big, int *__restrict__ big_count) {
    ClipV poly[8], tmp[8];
    poly[0].x = c0.x; poly[0].y = c0.y; poly[0].z = c0.z; poly[0].w = c0.w;
    poly[1].x = c1.x; poly[1].y = c1.y; poly[1].z = c1.z; poly[1].w = c1.w;
    poly[2].x = c2.x; poly[2].y = c2.y; poly[2].z = c2.z; poly[2].w = c2.w;
    int n = clip_poly(poly, 3, tmp, 0);
    if (n < 3) return;
    n = clip_poly(tmp, n, poly, 1);
    if (n < 3) return;
    int X[8], Y[8];
    double Z[8];
    for (int k = 0; k < n; k++) {
        double iw = __ddiv_rn(1.0, poly[k].w);
        double xn = __dmul_rn(poly[k].x, iw), yn = __dmul_rn(poly[k].y, iw), zn = __dmul_rn(poly[k].z, iw);
        X[k] = snap(__dmul_rn(__dadd_rn(xn, 1.0), __dmul_rn((double)W, 0.5)));
        Y[k] = snap(__dmul_rn(__dadd_rn(yn, 1.0), __dmul_rn((double)H, 0.5)));
        Z[k] = __dmul_rn(__dadd_rn(zn, 1.0), 0.5);
    }
    for (int k = 1; k + 1 < n; k++) {
        TriSetup s;
        s.X0 = X[0]; s.Y0 = Y[0]; s.z0 = Z[0];
        s.X1 = X[k]; s.Y1 = Y[k]; s.z1 = Z[k];
        s.X2 = X[k + 1]; s.Y2 = Y[k + 1]; s.z2 = Z[k + 1];
        s.mesh = mesh; s.pad = 0;
        TriRaster t;
        if (tri_prepare(s, H, W, t)) big[atomicAdd(big_count, 1)] = s;
    }
}

// Vertex stage + clipping + snapping of one triangle.  Returns the pixel-box area of the triangle if it joins
// the warp's flat list (t is then valid), 0 otherwise (rejected, or queued for k_big).
__device__ __forceinline__ int tri_setup(const MeshDesc &md, int mesh, const int *__restrict__ tri,
                                         const float *__restrict__ V, int H, int W, TriRaster &t,
                                         TriSetup *__restrict__ big, int *__restrict__ big_count) {
    float cv[3][4];
    bool inside_all = true;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        int vi = tri[k];
        if (vi < 0 || vi >= md.nv) return 0;
        const float *p = V + 3 * (md.v_begin + vi);
        const float x = p[0], y = p[1], z = p[2];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            float s = __fmul_rn(md.mvp[0 * 4 + r], x);
            s = __fadd_rn(s, __fmul_rn(md.mvp[1 * 4 + r], y));
            s = __fadd_rn(s, __fmul_rn(md.mvp[2 * 4 + r], z));
            s = __fadd_rn(s, md.mvp[3 * 4 + r]);
            cv[k][r] = s;
        }
        // the two clip tests of clip_poly: a triangle that passes both for every vertex comes out unchanged
        const double w = cv[k][3], zc = cv[k][2];
        inside_all = inside_all && __dadd_rn(w, zc) >= 0.0 && __dsub_rn(w, zc) >= 0.0;
    }
    if (!inside_all) {
        tri_setup_clipped(make_float4(cv[0][0], cv[0][1], cv[0][2], cv[0][3]), make_float4(cv[1][0], cv[1][1], cv[1][2], cv[1][3]),
                          make_float4(cv[2][0], cv[2][1], cv[2][2], cv[2][3]), mesh, H, W, big, big_count);
        return 0;
    }
    TriSetup s;
    {
        int X[3], Y[3];
        double Z[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const double iw = __ddiv_rn(1.0, (double)cv[k][3]);
            const double xn = __dmul_rn((double)cv[k][0], iw), yn = __dmul_rn((double)cv[k][1], iw),
                         zn = __dmul_rn((double)cv[k][2], iw);
            X[k] = snap(__dmul_rn(__dadd_rn(xn, 1.0), __dmul_rn((double)W, 0.5)));
            Y[k] = snap(__dmul_rn(__dadd_rn(yn, 1.0), __dmul_rn((double)H, 0.5)));
            Z[k] = __dmul_rn(__dadd_rn(zn, 1.0), 0.5);
        }
        s.X0 = X[0]; s.Y0 = Y[0]; s.z0 = Z[0];
        s.X1 = X[1]; s.Y1 = Y[1]; s.z1 = Z[1];
        s.X2 = X[2]; s.Y2 = Y[2]; s.z2 = Z[2];
        s.mesh = mesh; s.pad = 0;
    }
    if (!tri_prepare(s, H, W, t)) return 0;
    const int64_t box = (int64_t)(t.i1 - t.i0 + 1) * (t.j1 - t.j0 + 1);
    const int cmax = max(max(max(abs(s.X0), abs(s.X1)), max(abs(s.X2), abs(s.Y0))), max(abs(s.Y1), abs(s.Y2)));
    if (box > kFlatMax || cmax >= kCoordSafe || H > 8192 || W > 8192 || mesh >= (1 << 28)) {
        big[atomicAdd(big_count, 1)] = s;  // close-ups, far-off vertices
        return 0;
    }
    return (int)box;
}

// the list entry of a triangle: edge functions at the box's first pixel centre, biased by the top-left rule;
// per-pixel steps; depth constants
__device__ __forceinline__ void fill_rec(const TriRaster &t, int mesh, int start, TriRec &rec) {
    long long E0, E1, E2;
    edge_at(t, (long long)t.i0 * kSub + 128, (long long)t.j0 * kSub + 128, E0, E1, E2);
    rec.C0 = E0 + (t.tl0 ? 0 : -1);
    rec.C1 = E1 + (t.tl1 ? 0 : -1);
    rec.C2 = E2 + (t.tl2 ? 0 : -1);
    rec.A0 = (int)(-t.dy0 * kSub); rec.B0 = (int)(t.dx0 * kSub);
    rec.A1 = (int)(-t.dy1 * kSub); rec.B1 = (int)(t.dx1 * kSub);
    rec.A2 = (int)(-t.dy2 * kSub); rec.B2 = (int)(t.dx2 * kSub);
    rec.i0 = t.i0; rec.j0 = t.j0;
    rec.bw = t.i1 - t.i0 + 1;
    rec.magic = rec.bw > 1 ? 0xffffffffu / (unsigned)rec.bw + 1u : 0u;  // = ceil(2^32 / bw) for bw >= 2
    rec.mesh_tl = (mesh << 3) | (t.tl0 ? 1 : 0) | (t.tl1 ? 2 : 0) | (t.tl2 ? 4 : 0);
    rec.z0 = t.z0; rec.z1 = t.z1; rec.z2 = t.z2; rec.a2 = t.a2;
    rec.c = __ddiv_rn((double)kZMax, t.a2);
    rec.start = start;
}

// Depth of a covered pixel.  Same value as shade_pixel, bit for bit: the depth is
// q24 = rint(rn(rn(num / a2) * (2^24-1))); num * c with c = (2^24-1)/a2 differs from that product by < 1e-6
// (four roundings of relative 2^-53, guarded by |.| < 1e9), so unless it lies within 1e-6 of a rounding
// boundary (k + 1/2) both round to the same integer — and then the division is done after all.
__device__ __forceinline__ void shade_item(const CovItem &it, const TriRec &r, int H, int W, unsigned *__restrict__ zbuf) {
    const long long E0 = it.E0 + ((r.mesh_tl & 1) ? 0 : 1), E1 = it.E1 + ((r.mesh_tl & 2) ? 0 : 1),
                    E2 = it.E2 + ((r.mesh_tl & 4) ? 0 : 1);
    const double num = __dadd_rn(__dadd_rn(__dmul_rn((double)E0, r.z0), __dmul_rn((double)E1, r.z1)),
                                 __dmul_rn((double)E2, r.z2));
    const double qf = __dmul_rn(num, r.c);
    long long qi;
    if (fabs(qf) < 1.0e9 && fabs(__dsub_rn(__dsub_rn(qf, floor(qf)), 0.5)) > 1.0e-6) {
        qi = __double2ll_rn(qf);
        if (qi < 0) qi = 0;
    } else {
        const double qd = __dmul_rn(__ddiv_rn(num, r.a2), (double)kZMax);
        qi = __double2ll_rn(qd);
        if (!(qd > 0.0)) qi = 0;
    }
    if (qi > (long long)kZMax) qi = kZMax;
    atomicMin(zbuf + (size_t)(r.mesh_tl >> 3) * H * W + it.pix, (unsigned)qi);  // GL_LESS
}

__global__ void __launch_bounds__(kTrisTpb, 1024 / kTrisTpb) k_tris(const MeshDesc *__restrict__ meshes, const int *__restrict__ tri_mesh_start,
                                                   int n_mesh, int64_t n_tri_total, const float *__restrict__ V,
                                                   const int *__restrict__ F, int H, int W, unsigned *__restrict__ zbuf,
                                                   TriSetup *__restrict__ big, int *__restrict__ big_count) {
    __shared__ TriRec recs_sh[kTrisTpb / 32][32];
    __shared__ CovItem queue_sh[kTrisTpb / 32][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    TriRec *recs = recs_sh[warp];
    CovItem *queue = queue_sh[warp];
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int area = 0, mesh = 0;
    TriRaster t;
    if (gid < n_tri_total) {
        // which mesh does this triangle belong to: binary search over the per-mesh triangle offsets
        int lo = 0, hi = n_mesh;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (tri_mesh_start[mid] <= gid) lo = mid; else hi = mid;
        }
        mesh = lo;
        const MeshDesc &md = meshes[lo];
        const int f = (int)(gid - tri_mesh_start[lo]);
        area = tri_setup(md, lo, F + 3 * (md.f_begin + f), V, H, W, t, big, big_count);
    }
    // the warp's flat list holds its non-empty triangles in lane order: the one of rank k owns the indices
    // [end_k - area_k, end_k), and the ends are strictly increasing
    int incl = area;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;
    const unsigned have = __ballot_sync(0xffffffffu, area > 0);
    if (area) fill_rec(t, mesh, incl - area, recs[__popc(have & ((1u << lane) - 1u))]);
    const int my_end = area ? incl : -1;
    __syncwarp();
    int kbase = 0;  // triangles that end at or before p0 (warp-uniform)
    int cnt = 0;    // covered pixels waiting in the queue (warp-uniform, < 32 between rounds)
    for (int p0 = 0; p0 < total; p0 += 32) {
        const int p = p0 + lane;
        // bit i: a triangle ends at p0 + i, i.e. flat index p0 + i is the next one's first pixel
        const unsigned rel = (unsigned)(my_end - p0);
        const unsigned endm = __reduce_or_sync(0xffffffffu, rel < 32u ? 1u << rel : 0u);
        const int k = kbase + __popc(endm & (0xffffffffu >> (31 - lane)));  // triangles with end <= p
        kbase += __popc(endm);
        bool cov = false;
        CovItem it;
        if (p < total) {
            const TriRec &r = recs[k];
            const int q = p - r.start;
            const int row = r.bw > 1 ? (int)__umulhi((unsigned)q, r.magic) : q;
            const int col = q - row * r.bw;
            it.E0 = r.C0 + (long long)r.A0 * col + (long long)r.B0 * row;
            it.E1 = r.C1 + (long long)r.A1 * col + (long long)r.B1 * row;
            it.E2 = r.C2 + (long long)r.A2 * col + (long long)r.B2 * row;
            cov = (it.E0 | it.E1 | it.E2) >= 0;
            it.k = k;
            it.pix = (r.j0 + row) * W + (r.i0 + col);
        }
        const unsigned m = __ballot_sync(0xffffffffu, cov);
        if (cov) queue[cnt + __popc(m & ((1u << lane) - 1u))] = it;
        cnt += __popc(m);
        __syncwarp();
        if (cnt >= 32) {
            cnt -= 32;
            const CovItem c = queue[cnt + lane];
            shade_item(c, recs[c.k], H, W, zbuf);
            __syncwarp();  // the slots are free again
        }
    }
    if (lane < cnt) {
        const CovItem c = queue[lane];
        shade_item(c, recs[c.k], H, W, zbuf);
    }
}

// queued triangles: one BLOCK per triangle, threads stride over the pixel box (close-ups: a thread — or a
// warp — per triangle would serialise on the few huge boxes)
__global__ void __launch_bounds__(256) k_big(const TriSetup *__restrict__ big, const int *__restrict__ big_count,
                                             int H, int W, unsigned *__restrict__ zbuf) {
    const int count = *big_count;
    for (int q = blockIdx.x; q < count; q += gridDim.x) {
        TriRaster t;
        const TriSetup s = big[q];
        if (!tri_prepare(s, H, W, t)) continue;
        unsigned *zb = zbuf + (size_t)s.mesh * H * W;
        const int bw = t.i1 - t.i0 + 1, bh = t.j1 - t.j0 + 1;
        for (int p = threadIdx.x; p < bw * bh; p += blockDim.x) {
            int i = t.i0 + p % bw, j = t.j0 + p / bw;
            long long E0, E1, E2;
            edge_at(t, (long long)i * kSub + 128, (long long)j * kSub + 128, E0, E1, E2);
            shade_pixel(t, i, j, E0, E1, E2, zb, W);
        }
    }
}

__global__ void __launch_bounds__(256) k_resolve(const unsigned *__restrict__ z, float *__restrict__ depth, int64_t n) {
    // q / (2^24-1) as float: the double product with the rounded reciprocal gives the same float as the exact
    // division for every one of the 2^24 inputs (checked exhaustively, tests/test_oracle_golden.py).
    // Four pixels per thread: 16-byte loads and stores (cudaMalloc'ed / torch buffers are 16-byte aligned; a
    // caller's odd device pointer takes the scalar path).
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    const bool vec = ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(depth)) & 15) == 0;
    if (vec && i + 3 < n) {
        const uint4 q = *reinterpret_cast<const uint4 *>(z + i);
        float4 d;
        d.x = (float)__dmul_rn((double)q.x, 1.0 / 16777215.0);
        d.y = (float)__dmul_rn((double)q.y, 1.0 / 16777215.0);
        d.z = (float)__dmul_rn((double)q.z, 1.0 / 16777215.0);
        d.w = (float)__dmul_rn((double)q.w, 1.0 / 16777215.0);
        *reinterpret_cast<float4 *>(depth + i) = d;
    } else {
        for (int64_t k = i; k < n && k < i + 4; k++) depth[k] = (float)__dmul_rn((double)z[k], 1.0 / 16777215.0);
    }
}

// Renderer::RenderEdge's full-screen pass (render/shaders/edge_detection.frag:38-76) on the integer z-buffer:
// linearise the 3x3 neighbourhood with the SHADER's z_near/z_far (0.05 / 2.0 in the reference,
// render/renderer.cpp:95-96), mean absolute difference of the four opposite pairs, soft threshold 0.05-0.10,
// 5-texel border and background = 0, unorm8 output.  Float ops unfused and in the oracle's order.
__device__ __forceinline__ float edge_linearize(unsigned q, float zn, float zf) {
    if (q == kZMax) return -1.0f;
    const float z = (float)__dmul_rn((double)q, 1.0 / 16777215.0);  // == q / (2^24-1) for every q (exhaustive)
    float a = __fmul_rn(2.0f, zn);
    a = __fmul_rn(a, zf);
    float b = __fmul_rn(2.0f, z);
    b = __fsub_rn(b, 1.0f);
    const float c = __fsub_rn(zf, zn);
    b = __fmul_rn(b, c);
    float d = __fadd_rn(zf, zn);
    d = __fsub_rn(d, b);
    return __fdiv_rn(a, d);
}

__global__ void __launch_bounds__(256) k_edge_mask(const unsigned *__restrict__ z, int n_mesh, int H, int W, float zn,
                                                   float zf, unsigned char *__restrict__ edge,
                                                   unsigned char *__restrict__ mask) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (int64_t)n_mesh * H * W) return;
    const int i = (int)(p % W), j = (int)((p / W) % H);
    const unsigned zc = z[p];
    // the reference clears the colour buffer to 1.0 and reads GL_RED back (render/renderer.cpp:411-422): background =
    // 255; its depth shader writes no colour, so covered pixels are undefined there — defined here as 0
    if (mask) mask[p] = zc != kZMax ? 0 : 255;
    if (!edge) return;
    unsigned char e = 0;
    if (i >= 5 && i < W - 5 && j >= 5 && j < H - 5 && zc != kZMax) {
        float v[9];
        int k = 0;
#pragma unroll
        for (int di = -1; di <= 1; di++)
#pragma unroll
            for (int dj = -1; dj <= 1; dj++) v[k++] = edge_linearize(z[p + (int64_t)dj * W + di], zn, zf);
        float s = fabsf(__fsub_rn(v[1], v[7]));
        s = __fadd_rn(s, fabsf(__fsub_rn(v[5], v[3])));
        s = __fadd_rn(s, fabsf(__fsub_rn(v[0], v[8])));
        s = __fadd_rn(s, fabsf(__fsub_rn(v[2], v[6])));
        const float delta = __fmul_rn(0.25f, s);
        float c;
        if (delta < 0.05f) c = 0.0f;
        else if (delta >= 0.10f) c = 1.0f;
        else c = __fdiv_rn(__fsub_rn(delta, 0.05f), __fsub_rn(0.10f, 0.05f));
        e = (unsigned char)__float2int_rn(__fmul_rn(c, 255.0f));
    }
    edge[p] = e;
}

// ---- host-side camera maths (float, evaluated in the order the reference / glm evaluate it) ---------
// Renderer::SetCamera(zn, zf, intrinsics): frustum extents with top/bottom flipped (render/renderer.cpp:259-267)
// fed to glm::frustum (RH, z in [-1,1]).
void projection_matrix(float zn, float zf, float fx, float fy, float cx, float cy, int H, int W, float P[16]) {
    volatile float left = -cx / fx * zn;
    volatile float right = (float)(((double)(float)W - 1.0 - (double)cx) / (double)fx * (double)zn);
    volatile float bottom = cy / fy * zn;
    volatile float top = (cy - (float)(H - 1)) / fy * zn;
    for (int i = 0; i < 16; i++) P[i] = 0.0f;
    P[0] = (2.0f * zn) / (right - left);
    P[5] = (2.0f * zn) / (top - bottom);
    P[8] = (right + left) / (right - left);
    P[9] = (top + bottom) / (top - bottom);
    P[10] = -(zf + zn) / (zf - zn);
    P[11] = -1.0f;
    P[14] = -(2.0f * zf * zn) / (zf - zn);
}

void mat4f_mul(const float a[16], const float b[16], float c[16]) {  // column-major, left-to-right sums
    float r[16];
    for (int col = 0; col < 4; col++)
        for (int row = 0; row < 4; row++) {
            volatile float s = a[0 * 4 + row] * b[col * 4 + 0];
            s = s + a[1 * 4 + row] * b[col * 4 + 1];
            s = s + a[2 * 4 + row] * b[col * 4 + 2];
            s = s + a[3 * 4 + row] * b[col * 4 + 3];
            r[col * 4 + row] = s;
        }
    memcpy(c, r, sizeof(r));
}

}  // namespace

}  // namespace vb

extern "C" int vb200_render_depth_batch(const float *V_concat, const int64_t *v_off, const int32_t *F_concat,
                                        const int64_t *f_off, int32_t n_mesh, const float *model_T,
                                        const float view_T[16], float zn, float zf, float fx, float fy, float cx,
                                        float cy, int H, int W, int device, uint32_t *out_z24, float *out_depth) {
    return vb200_render_depth_batch_ex(V_concat, v_off, F_concat, f_off, n_mesh, model_T, view_T, zn, zf, fx, fy, cx,
                                       cy, H, W, device, out_z24, out_depth, 0, nullptr);
}

static int render_batch_impl(const float *V_concat, const int64_t *v_off, const int32_t *F_concat,
                             const int64_t *f_off, int32_t n_mesh, const float *model_T, const float view_T[16],
                             float zn, float zf, float fx, float fy, float cx, float cy, int H, int W, int device,
                             uint32_t *out_z24, float *out_depth, uint8_t *out_edge, uint8_t *out_mask,
                             float edge_zn, float edge_zf, int outputs_on_device, float *kernel_ms);

extern "C" int vb200_render_depth_batch_ex(const float *V_concat, const int64_t *v_off, const int32_t *F_concat,
                                           const int64_t *f_off, int32_t n_mesh, const float *model_T,
                                           const float view_T[16], float zn, float zf, float fx, float fy,
                                           float cx, float cy, int H, int W, int device, uint32_t *out_z24,
                                           float *out_depth, int outputs_on_device, float *kernel_ms) {
    return render_batch_impl(V_concat, v_off, F_concat, f_off, n_mesh, model_T, view_T, zn, zf, fx, fy, cx, cy, H, W,
                             device, out_z24, out_depth, nullptr, nullptr, 0.05f, 2.0f, outputs_on_device, kernel_ms);
}

extern "C" int vb200_render_edge_mask_batch(const float *V_concat, const int64_t *v_off, const int32_t *F_concat,
                                            const int64_t *f_off, int32_t n_mesh, const float *model_T,
                                            const float view_T[16], float zn, float zf, float fx, float fy,
                                            float cx, float cy, int H, int W, int device, float edge_z_near,
                                            float edge_z_far, uint8_t *out_edge, uint8_t *out_mask,
                                            int outputs_on_device) {
    return render_batch_impl(V_concat, v_off, F_concat, f_off, n_mesh, model_T, view_T, zn, zf, fx, fy, cx, cy, H, W,
                             device, nullptr, nullptr, out_edge, out_mask, edge_z_near, edge_z_far, outputs_on_device,
                             nullptr);
}

static int render_batch_impl(const float *V_concat, const int64_t *v_off, const int32_t *F_concat,
                             const int64_t *f_off, int32_t n_mesh, const float *model_T, const float view_T[16],
                             float zn, float zf, float fx, float fy, float cx, float cy, int H, int W, int device,
                             uint32_t *out_z24, float *out_depth, uint8_t *out_edge, uint8_t *out_mask,
                             float edge_zn, float edge_zf, int outputs_on_device, float *kernel_ms) {
    using namespace vb;
    if (n_mesh < 0 || H <= 0 || W <= 0 || !v_off || !f_off || !view_T || (n_mesh > 0 && !model_T))
        return VB200_ERR_INVALID;
    if (n_mesh == 0) return VB200_OK;
    const int64_t nv = v_off[n_mesh] - v_off[0], nf = f_off[n_mesh] - f_off[0];
    if (nv < 0 || nf < 0 || (nv > 0 && !V_concat) || (nf > 0 && !F_concat)) return VB200_ERR_INVALID;
    VB_TRY(select_device(device));
    // camera: projection (SetCamera(zn,zf,...)) and view = diag(1,-1,-1,1) * pose (SetCamera(pose),
    // render/renderer.cpp:284-300); MVP = (P*V)*M as the vertex shader's left-to-right product
    float P[16], Vw[16], PV[16];
    projection_matrix(zn, zf, fx, fy, cx, cy, H, W, P);
    const float v2g[16] = {1, 0, 0, 0, 0, -1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1};
    mat4f_mul(v2g, view_T, Vw);
    mat4f_mul(P, Vw, PV);
    std::vector<MeshDesc> meshes((size_t)n_mesh);
    std::vector<int> tri_start((size_t)n_mesh + 1);
    for (int m = 0; m < n_mesh; m++) {
        MeshDesc &d = meshes[m];
        mat4f_mul(PV, model_T + 16 * (size_t)m, d.mvp);
        d.v_begin = v_off[m] - v_off[0];
        d.f_begin = f_off[m] - f_off[0];
        int64_t mv = v_off[m + 1] - v_off[m], mf = f_off[m + 1] - f_off[m];
        if (mv < 0 || mf < 0 || mv > 0x7fffffff || f_off[m + 1] - f_off[0] > 0x7fffffff) return VB200_ERR_INVALID;
        d.nv = (int)mv;
        d.nf = (int)mf;
        tri_start[m] = (int)d.f_begin;
    }
    tri_start[n_mesh] = (int)nf;

    cudaStream_t st;
    VB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};
    const int64_t npix = (int64_t)n_mesh * H * W;
    // stream-ordered allocations: with the pool's release threshold raised (select_device) a caller that renders
    // batch after batch reuses the same 2 x 157 MB instead of paying cudaMalloc + cudaFree (device-wide
    // synchronisations) on every call
    DevBuf<float> d_V(st), d_depth(st);
    DevBuf<int> d_F(st), d_tri_start(st), d_big_count(st);
    DevBuf<MeshDesc> d_mesh(st);
    DevBuf<unsigned> d_z(st);
    DevBuf<TriSetup> d_big(st);
    VB_CUDA(d_V.alloc(3 * (size_t)std::max<int64_t>(nv, 1)));
    VB_CUDA(d_F.alloc(3 * (size_t)std::max<int64_t>(nf, 1)));
    VB_CUDA(d_mesh.alloc((size_t)n_mesh));
    VB_CUDA(d_tri_start.alloc((size_t)n_mesh + 1));
    if (!(outputs_on_device && out_z24)) VB_CUDA(d_z.alloc((size_t)npix));
    VB_CUDA(d_big.alloc(3 * (size_t)std::max<int64_t>(nf, 1)));
    VB_CUDA(d_big_count.alloc(1));
    if (nv) VB_CUDA(cudaMemcpyAsync(d_V.p, V_concat + 3 * v_off[0], sizeof(float) * 3 * (size_t)nv, cudaMemcpyHostToDevice, st));
    if (nf) VB_CUDA(cudaMemcpyAsync(d_F.p, F_concat + 3 * f_off[0], sizeof(int) * 3 * (size_t)nf, cudaMemcpyHostToDevice, st));
    VB_CUDA(cudaMemcpyAsync(d_mesh.p, meshes.data(), sizeof(MeshDesc) * (size_t)n_mesh, cudaMemcpyHostToDevice, st));
    VB_CUDA(cudaMemcpyAsync(d_tri_start.p, tri_start.data(), sizeof(int) * ((size_t)n_mesh + 1), cudaMemcpyHostToDevice, st));
    VB_CUDA(cudaMemsetAsync(d_big_count.p, 0, sizeof(int), st));
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    struct EventGuard { cudaEvent_t &a, &b; ~EventGuard() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); } } eguard{ev0, ev1};
    if (kernel_ms) {
        VB_CUDA(cudaEventCreate(&ev0));
        VB_CUDA(cudaEventCreate(&ev1));
        VB_CUDA(cudaEventRecord(ev0, st));
    }
    // with device outputs the kernels write straight into the caller's buffers
    unsigned *zbuf = (outputs_on_device && out_z24) ? out_z24 : d_z.p;
    k_clear<<<kNumSMsB200 * 8, 256, 0, st>>>(zbuf, npix);
    if (nf) {
        k_tris<<<div_up(nf, kTrisTpb), kTrisTpb, 0, st>>>(d_mesh.p, d_tri_start.p, n_mesh, nf, d_V.p, d_F.p, H, W, zbuf, d_big.p, d_big_count.p);
        k_big<<<kNumSMsB200 * 8, 256, 0, st>>>(d_big.p, d_big_count.p, H, W, zbuf);
    }
    VB_CUDA(cudaGetLastError());
    if (out_depth) {
        float *dd = out_depth;
        if (!outputs_on_device) {
            VB_CUDA(d_depth.alloc((size_t)npix));
            dd = d_depth.p;
        }
        k_resolve<<<div_up(div_up(npix, 4), 256), 256, 0, st>>>(zbuf, dd, npix);
        VB_CUDA(cudaGetLastError());
    }
    DevBuf<unsigned char> d_edge(st), d_mask(st);
    if (out_edge || out_mask) {
        unsigned char *de = out_edge, *dm = out_mask;
        if (!outputs_on_device) {
            if (out_edge) { VB_CUDA(d_edge.alloc((size_t)npix)); de = d_edge.p; }
            if (out_mask) { VB_CUDA(d_mask.alloc((size_t)npix)); dm = d_mask.p; }
        }
        k_edge_mask<<<div_up(npix, 256), 256, 0, st>>>(zbuf, n_mesh, H, W, edge_zn, edge_zf, de, dm);
        VB_CUDA(cudaGetLastError());
        if (!outputs_on_device) {
            if (out_edge) VB_CUDA(cudaMemcpyAsync(out_edge, d_edge.p, (size_t)npix, cudaMemcpyDeviceToHost, st));
            if (out_mask) VB_CUDA(cudaMemcpyAsync(out_mask, d_mask.p, (size_t)npix, cudaMemcpyDeviceToHost, st));
        }
    }
    if (kernel_ms) VB_CUDA(cudaEventRecord(ev1, st));
    if (!outputs_on_device) {
        if (out_depth) VB_CUDA(cudaMemcpyAsync(out_depth, d_depth.p, sizeof(float) * (size_t)npix, cudaMemcpyDeviceToHost, st));
        if (out_z24) VB_CUDA(cudaMemcpyAsync(out_z24, d_z.p, sizeof(unsigned) * (size_t)npix, cudaMemcpyDeviceToHost, st));
    }
    VB_CUDA(cudaStreamSynchronize(st));
    if (kernel_ms) VB_CUDA(cudaEventElapsedTime(kernel_ms, ev0, ev1));
    return VB200_OK;
}
