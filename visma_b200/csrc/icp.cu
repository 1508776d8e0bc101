// icp.cu — the batched ICP operator: correspondence search fused with the estimator's reductions, the
// small solves, and the convergence loop, all resident on the GPU.
//
// Reference path replaced (O3D = thirdparty/Open3D):
//   open3d::RegistrationICP                         O3D/src/Core/Registration/Registration.cpp:141-186
//   GetRegistrationResultAndCorrespondences         Registration.cpp:41-96      } k_pass: one launch does
//   PointCloud::Transform                           Geometry/PointCloud.cpp:75-87 } transform + 1-NN + the
//   TransformationEstimationPointToPlane rows       TransformationEstimation.cpp:82-90 } per-correspondence
//   ComputeJTJandJTr                                Utility/Eigen.cpp:137-182   } products + block reduction
//   TransformationEstimationPointToPoint / cicp     TransformationEstimation.cpp:47-59, src/constrained_ICP.cpp:25-37
//   SolveJacobianSystemAndObtainExtrinsicMatrix     Utility/Eigen.cpp:88-106    } k_solve: one thread per
//   Eigen::umeyama                                  3rdparty/Eigen/.../Umeyama.h:93-162 } problem, fixed-order reduce
//   feh::RegisterModelToScene                       src/annotation.cpp:29-64
//
// Differences from the reference that are by design:
//   - the accumulated transform is applied to the ORIGINAL source points each iteration instead of
//     transforming the cloud incrementally (Registration.cpp:175): one read, no write, no drift;
//   - correspondences are never materialised between the search and the estimator (the reference writes a
//     CorrespondenceSet and re-gathers it): the matched target point/normal are consumed in registers;
//   - reductions have a fixed order (warp halving tree -> warps -> blocks), so results are deterministic;
//     the reference's OpenMP merge order is thread-arrival order.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "linalg.cuh"
#include "scene.cuh"
#include "sort.cuh"

namespace vb {

namespace {

// Part B of the pass runs over a worklist of the blocks that listed anything (k_pass_b_wl) unless the older
// one-block-per-(block, warp) launch is asked for (-DVB_PB_PER_BLOCK; the fused iteration tail needs it too).
#if !defined(VB_PB_PER_BLOCK) && !defined(VB_FUSE_SOLVE) && !defined(VB_PB_WORKLIST)
#define VB_PB_WORKLIST
#endif
#ifndef VB_PASS_TPB
#define VB_PASS_TPB 128
#endif
#ifndef VB_PASS_MINBLOCKS
#define VB_PASS_MINBLOCKS (768 / VB_PASS_TPB)  // 24 resident warps per SM -> 80 registers per thread
#endif
constexpr int kPassTpb = VB_PASS_TPB;
constexpr int kPassWarps = kPassTpb / 32;  // every warp writes its own partial: no block-level barrier
#ifndef VB_PASS_PTS
#define VB_PASS_PTS 4
#endif
constexpr int kPtsPerThread = VB_PASS_PTS;
constexpr int kChunk = kPassTpb * kPtsPerThread;  // source points per block
constexpr int kAcc = 32;                          // accumulator slots (padded)
constexpr int kBuckets = 32768;                   // spatial buckets per source cloud (15-bit key)

// accumulator slots, MODE 1 (point-to-plane): 0..20 JTJ upper triangle row-major, 21..26 JTr
// MODE 0 (point-to-point): 0..2 sum s', 3..5 sum d', 6..14 sum d' s'^T (row-major), 15 sum |s'|^2
// both: 30 = sum d2 (exact NN distances), 31 = count
constexpr int kSlotD2 = 30, kSlotCount = 31;

struct ProbState {
    double T[16];
    double fitness, rmse, prev_fitness, prev_rmse;
    int ncorr, iters, done, pad;
};

struct BlockTask {
    int prob;       // problem index
    int src_begin;  // first source point (global sorted position)
    int count;      // points in this block's chunk
    int corr_begin; // where this chunk's matches go in corr_j
};

struct ProbDesc {
    int cloud;
    int npts;
    int blk_begin, blk_count;
    int src_begin;
    int corr_begin;
};

struct PassParams {
    double r2;      // (double)(float)(max_dist^2)  (KDTreeFlann.cpp:185)
    float r2_ub;    // f32 upper bound of r2 including the screening band
    float slack;    // how far beyond the answer's reach a search looks, so that its result survives small moves
    float pos_err;  // bound on the error of a distance between two centred-f32 query positions
    int use_cache;  // 0: every point is searched in every pass (dev knob VB200_NN_CACHE=0: the ablation bench.py reports)
#ifdef VB_PB_WORKLIST
    int *work;      // part-A blocks that listed anything in this pass, in order of arrival
    int *work_ctr;  // [parity][2]: {blocks listed, part-B items handed out}; passes alternate between the two pairs
    int parity;
#endif
};
#if defined(VB_PB_WORKLIST) && defined(VB_FUSE_SOLVE)
#error "the worklist part B has no fused iteration tail"
#endif

#ifndef VB_SLACK_PCT
#define VB_SLACK_PCT 8
#endif
inline PassParams make_pass_params(const GridParams &g, double max_dist) {
    PassParams pp;
    pp.r2 = (double)(float)(max_dist * max_dist);
    pp.r2_ub = r2_upper_bound(g, pp.r2);
    pp.slack = g.fine * (VB_SLACK_PCT * 0.01f);
    pp.pos_err = g.band_a;  // band_a = 2 sqrt(3) e with e the per-axis error of such a difference: a 2x margin
    const char *nc = getenv("VB200_NN_CACHE");
    pp.use_cache = !(nc && nc[0] == '0');
    return pp;
}

struct SolveParams {
    double rel_fitness, rel_rmse;
    double g[3];
    int max_iter;
    int estimator;
    int *ndone;  // nullable: counts the problems that have finished (lets the host stop enqueueing iterations)
};

// ---- warp-level vector reduction: 32 slots over 32 lanes in 31 exchange steps (recursive halving).
// After the call lane L holds the warp total of slot L.  Fixed order => deterministic.
__device__ __forceinline__ double warp_reduce_slots(double (&v)[kAcc]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) {
        const bool up = (lane & h) != 0;
#pragma unroll
        for (int j = 0; j < h; j++) {
            double send = up ? v[j] : v[j + h];
            double keep = up ? v[j + h] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
    return v[0];
}

template <int MODE>
__device__ __forceinline__ void contributions(bool matched, double d2, const double *vs, const double *vt,
                                              const double *nt, const double *cref, double (&v)[kAcc]) {
#pragma unroll
    for (int j = 0; j < kAcc; j++) v[j] = 0.0;
    if (!matched) return;
    if (MODE == 1) {
        // r = (vs - vt).nt ; J = [vs x nt ; nt]  (TransformationEstimation.cpp:87-89)
        double r = (vs[0] - vt[0]) * nt[0] + (vs[1] - vt[1]) * nt[1] + (vs[2] - vt[2]) * nt[2];
        double J[6] = {vs[1] * nt[2] - vs[2] * nt[1], vs[2] * nt[0] - vs[0] * nt[2],
                       vs[0] * nt[1] - vs[1] * nt[0], nt[0], nt[1], nt[2]};
        int k = 0;
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
            for (int b = a; b < 6; b++) v[k++] = J[a] * J[b];
#pragma unroll
        for (int a = 0; a < 6; a++) v[21 + a] = J[a] * r;
    } else {
        // moments of (s', d') = (vs - c, vt - c): enough for Eigen::umeyama's means, Sigma and src_var
        double s[3] = {vs[0] - cref[0], vs[1] - cref[1], vs[2] - cref[2]};
        double d[3] = {vt[0] - cref[0], vt[1] - cref[1], vt[2] - cref[2]};
#pragma unroll
        for (int a = 0; a < 3; a++) { v[a] = s[a]; v[3 + a] = d[a]; }
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) v[6 + 3 * a + b] = d[a] * s[b];
        v[15] = s[0] * s[0] + s[1] * s[1] + s[2] * s[2];
    }
    v[kSlotD2] = d2;
    v[kSlotCount] = 1.0;
}

// per-warp shared scratch: the search's run lists, then the estimator rows (never live at the same time)
constexpr int kRowStride = 10;  // doubles per staged row: 80 B keeps the 16-byte row stores conflict-free
struct __align__(16) WarpScratch {
    union {
        LaneRuns<32> runs;
        double rows[32 * kRowStride];
    };
};
constexpr int kPart = 64;  // doubles per warp partial: the 8x8 Gram matrix of the staged rows
static_assert(32 * kPtsPerThread <= 256, "hard_ids holds a warp's point ids in one byte");
static_assert(kPassTpb == kPassWarps * 32 && kPassWarps * (kPart / 2) == kPassTpb, "k_pass_a zeroes part B's rows with one store per thread");

// D(8x8) += A(8x4) * B(4x8) in FP64 on the tensor cores (DMMA).  Lane l holds A[l>>2][l&3], B[l&3][l>>2] and
// D[l>>2][2(l&3) + {0,1}].
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// what a source point remembers from its last search (16 B, same indexing as corr_s): the query position then
// (centred f32, as QueryCtx) and `sec`, a lower bound of the true squared distance from there to every target
// point other than its match.  memset 0xff = NaN = "knows nothing".
struct __align__(16) NNCache {
    float qx, qy, qz, sec;
};

// One ICP correspondence pass for every active problem: transform + radius-bounded 1-NN + estimator
// products + reduction.  grid = one block per kChunk source points of one problem; every warp owns
// kPtsPerThread batches of 32 consecutive (spatially sorted) points.
//
// 1. Cached-neighbour test (Greenspan & Godin 2001, exact): a point whose last search left the bound `sec`
//    and which has since moved by delta still has the same nearest neighbour b if
//    |q - b| < sqrt(sec) - delta (triangle inequality, all f32 error bands on the safe side).  Such a point
//    needs no search at all: its row goes straight into the estimator sums.  Once an alignment settles this
//    is nearly every point, and the pass becomes a streaming gather/reduce.
// 2. The points that fail the test are compacted per warp (ballot ranks, deterministic) and searched 32 at a
//    time with all lanes busy (nn_search_hybrid); each search refreshes the point's cache.
// The decision is always taken in double from the exact coordinates: d2 = |q - b|^2 with FLANN's operation
// order, accepted iff d2 < (double)(float)(r*r).
//
// Estimator sums: every matched lane stages one row x of 8 doubles — point-to-plane [J (6), r, 0] with
// r = (vs - vt).nt, J = [vs x nt ; nt] (TransformationEstimation.cpp:87-89); point-to-point
// [s' (3), d' (3), 1, 0] with (s', d') = (vs - c, vt - c) — and the warp accumulates the Gram matrix
// sum_i x_i x_i^T with 8 DMMA instructions per 32 rows: JTJ = D[0..5][0..5], JTr = D[0..5][6]
// (resp. the umeyama moments sum d' s'^T = D[3..5][0..2], sum s' = D[0..2][6], sum d' = D[3..5][6]).  The
// accumulator fragment lives in two registers per lane for the whole warp.  (The first version formed 27
// products per lane and reduced 32 slots with a 31-step shuffle tree: ~280 instructions and 64 live
// registers per batch.)
template <int MODE>
struct PassCtx {
    const GridDev &G;
    const double *T;  // the problem's current transform (global memory, warp-uniform loads)
    double c0 = 0.0, c1 = 0.0;  // this lane's two entries of the warp's Gram matrix
    double sum_d2 = 0.0;        // exact NN distances (Registration.cpp:68), this lane's share
    int count = 0;

    __device__ __forceinline__ PassCtx(const GridDev &g, const double *t) : G(g), T(t) {}

    __device__ __forceinline__ void transform(const double *__restrict__ p, double (&vs)[3]) const {
        const double px = p[0], py = p[1], pz = p[2];
        vs[0] = T[0] * px + T[1] * py + T[2] * pz + T[3];
        vs[1] = T[4] * px + T[5] * py + T[6] * pz + T[7];
        vs[2] = T[8] * px + T[9] * py + T[10] * pz + T[11];
    }

    // the exact decision for match candidate bs (>= 0) and, if accepted, the lane's estimator row
    __device__ __forceinline__ bool row_of(int bs, const double (&vs)[3], double r2, double (&x)[8]) {
        const double *t = G.xyz + kPtStride * (int64_t)bs;
        const double vt[3] = {t[0], t[1], t[2]};
        const double d2 = l2_exact(vs[0], vs[1], vs[2], vt);
        return row_known(bs, vt, d2, vs, r2, x);
    }

    // the same from a target point already loaded and its exact distance
    // (nt_loaded: the target normal when the caller has fetched it already, alongside the point)
    __device__ __forceinline__ bool row_known(int bs, const double (&vt)[3], double d2, const double (&vs)[3], double r2,
                                              double (&x)[8], const double *nt_loaded = nullptr) {
        if (!(d2 < r2)) return false;
        if (MODE == 1) {
            const double *nn = nt_loaded ? nt_loaded : G.nrm + kPtStride * (int64_t)bs;
            const double nt[3] = {nn[0], nn[1], nn[2]};
            x[0] = vs[1] * nt[2] - vs[2] * nt[1];
            x[1] = vs[2] * nt[0] - vs[0] * nt[2];
            x[2] = vs[0] * nt[1] - vs[1] * nt[0];
            x[3] = nt[0]; x[4] = nt[1]; x[5] = nt[2];
            x[6] = (vs[0] - vt[0]) * nt[0] + (vs[1] - vt[1]) * nt[1] + (vs[2] - vt[2]) * nt[2];
        } else {
            // cref = translation part of T: keeps the moments O(object size) instead of O(scene size)
            const double cr[3] = {T[3], T[7], T[11]};
            x[0] = vs[0] - cr[0]; x[1] = vs[1] - cr[1]; x[2] = vs[2] - cr[2];
            x[3] = vt[0] - cr[0]; x[4] = vt[1] - cr[1]; x[5] = vt[2] - cr[2];
            x[6] = 1.0;
        }
        sum_d2 += d2;
        count += 1;
        return true;
    }

    // all 32 lanes: stage the rows (zeros for lanes without a match) and accumulate their Gram matrix
    __device__ __forceinline__ void accumulate(double *rows, const double (&x)[8]) {
        const int lane = threadIdx.x & 31;
        __syncwarp();  // the run lists are dead: the scratch now holds the estimator rows
        double2 *row = reinterpret_cast<double2 *>(rows + lane * kRowStride);
        row[0] = make_double2(x[0], x[1]);
        row[1] = make_double2(x[2], x[3]);
        row[2] = make_double2(x[4], x[5]);
        row[3] = make_double2(x[6], x[7]);
        __syncwarp();
#pragma unroll
        for (int ch = 0; ch < 8; ch++) {
            const double a = rows[(4 * ch + (lane & 3)) * kRowStride + (lane >> 2)];
            dmma_m8n8k4(c0, c1, a, a);
        }
        __syncwarp();  // rows consumed before anything else reuses the scratch
    }

    // the warp's partial: D[l>>2][2(l&3) + {0,1}] = entries 2l, 2l+1 of the row-major 8x8, one coalesced
    // 512-byte row.  Row 7 of D is identically zero (x[7] = 0); its last two entries carry the inlier count
    // and the sum of exact NN distances (fixed-order butterfly: deterministic).
    __device__ __forceinline__ double2 lane_pair() {
        const int lane = threadIdx.x & 31;
        double dd = sum_d2;
        int cnt = count;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            dd += __shfl_xor_sync(0xffffffffu, dd, o);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        double a = c0, b = c1;
        if (lane == 31) { a = (double)cnt; b = dd; }
        return make_double2(a, b);
    }
    __device__ __forceinline__ void write_partial(double *__restrict__ out_row) {
        reinterpret_cast<double2 *>(out_row)[threadIdx.x & 31] = lane_pair();
    }
};

#ifndef VB_COOP_WARP_LANES
#define VB_COOP_WARP_LANES 12  // more lanes than this with a long reach: the whole warp takes the shared walk
#endif
#ifndef VB_PASS_A_MINBLOCKS
#define VB_PASS_A_MINBLOCKS 12
#endif
constexpr int kRowsPerBlock = 1 + kPassWarps;  // partial rows per block: k_pass_a's (block total), then k_pass_b's warps

// ---- pass, part A: every point.  Streaming: transform, cached-neighbour test, and for the points it settles
// the exact decision and the estimator row.  The rest are listed per warp (ballot ranks: deterministic order)
// for part B.  No search code in here, so the kernel runs at high occupancy: it is a gather/reduce bound by
// memory latency and bandwidth.
template <int MODE>
__global__ void __launch_bounds__(kPassTpb, VB_PASS_A_MINBLOCKS) k_pass_a(
    GridDev G, const double *__restrict__ src_xyz, const BlockTask *__restrict__ tasks,
    const ProbState *__restrict__ states, double *__restrict__ partials, int *__restrict__ corr_s,
    const NNCache *__restrict__ cache, int *__restrict__ second_s, unsigned char *__restrict__ hard_ids,
    int *__restrict__ hard_cnt, PassParams pp) {
    const BlockTask task = tasks[blockIdx.x];
    const ProbState *st = states + task.prob;
#ifdef VB_PB_WORKLIST
    // the counters the NEXT pass will use: idle since the previous pass's part B finished
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        pp.work_ctr[2 * (pp.parity ^ 1)] = 0;
        pp.work_ctr[2 * (pp.parity ^ 1) + 1] = 0;
    }
#endif
    if (st->done) return;
    __shared__ __align__(16) double rows_sh[kPassWarps][32 * kRowStride];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    PassCtx<MODE> ctx(G, st->T);
    unsigned char *my_hard = hard_ids + ((int64_t)blockIdx.x * kPassWarps + warp) * (32 * kPtsPerThread);
    int nhard = 0;  // warp-uniform
#pragma unroll 1
    for (int k = 0; k < kPtsPerThread; k++) {
        const int local = (warp * kPtsPerThread + k) * 32 + lane;
        const bool valid = local < task.count;
        double x[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        bool hard = false;
        if (valid) {
            double vs[3];
            ctx.transform(src_xyz + 3 * (int64_t)(task.src_begin + local), vs);
            QueryCtx c;
            const bool inside = make_query(G.p, vs[0], vs[1], vs[2], c);
            const int slot = task.corr_begin + local;
            if (!inside) {
                corr_s[slot] = -1;  // farther than a cell outside the grid: no neighbour within the radius
            } else {
                hard = true;
#if defined(VB_PA_NOHI) && !defined(VB_NO_NN_CACHE)
                // The cached-neighbour test without the f32 screening array: how far the point has moved since
                // its last search is known from the streamed cache entry alone, and unless that is less than
                // the entry's bound nothing below can succeed — no gather at all for such a point (every point,
                // in an alignment's first iterations).  Otherwise the match's exact coordinates are needed
                // anyway: |q - b| comes from them in double (rounded up to f32: tighter than the screening
                // distance + band), one level of dependent loads fewer and one 32-byte sector per point less.
                const int prior = pp.use_cache ? corr_s[slot] : -1;
                if (prior >= 0) {
                    const NNCache m = cache[slot];
                    const float mx = c.qx - m.qx, my = c.qy - m.qy, mz = c.qz - m.qz;
                    const float mv = sqrtf(fmaf(mz, mz, fmaf(my, my, mx * mx))) * 1.000001f + pp.pos_err;
                    const float lim = sqrtf(fabsf(m.sec)) * 0.999999f;  // NaN (no cache) compares false
                    if (mv < lim) {
                        const double *tp = G.xyz + kPtStride * (int64_t)prior;
                        const double vt[3] = {tp[0], tp[1], tp[2]};
                        // (the normal is fetched with the point, not after the test: a point that gets here
                        // nearly always passes, and the two gathers then overlap)
                        double nt[3] = {0.0, 0.0, 0.0};
                        if (MODE == 1 && m.sec >= 0.0f) {
                            const double *np = G.nrm + kPtStride * (int64_t)prior;
                            nt[0] = np[0]; nt[1] = np[1]; nt[2] = np[2];
                        }
                        const double da = l2_exact(vs[0], vs[1], vs[2], vt);
                        if (m.sec >= 0.0f) {
                            // |q - b| (upper bound) + move (upper bound) < distance to anything else then (lower bound)
                            if (sqrtf(__double2float_ru(da)) * 1.000001f + mv < lim) {
                                hard = false;
                                // still the nearest, but it may have left the radius: then there is no
                                // correspondence (and nothing to remember: corr_s doubles as the list)
                                if (!ctx.row_known(prior, vt, da, vs, pp.r2, x, MODE == 1 ? nt : nullptr)) corr_s[slot] = -1;
                            }
                        } else {
                            // a two-candidate entry: if both are still nearer than anything else can be, the
                            // winner between them is taken in double
                            const int other = second_s[slot];
                            if (other >= 0) {
                                const double *up = G.xyz + kPtStride * (int64_t)other;
                                const double vu[3] = {up[0], up[1], up[2]};
                                const double db = l2_exact(vs[0], vs[1], vs[2], vu);
                                if (sqrtf(__double2float_ru(fmax(da, db))) * 1.000001f + mv < lim) {
                                    hard = false;
                                    const bool first = da < db || (da == db && __ldg(G.orig + prior) < __ldg(G.orig + other));
                                    if (!first) { corr_s[slot] = other; second_s[slot] = prior; }
                                    const bool ok = first ? ctx.row_known(prior, vt, da, vs, pp.r2, x)
                                                          : ctx.row_known(other, vu, db, vs, pp.r2, x);
                                    if (!ok) corr_s[slot] = -1;
                                }
                            }
                        }
                    }
                }
#elif !defined(VB_NO_NN_CACHE)
                const int prior = pp.use_cache ? corr_s[slot] : -1;
#ifdef VB_PA_PRETEST
                // how far the point has moved since its last search is known from the streamed cache entry alone:
                // unless that is less than the entry's bound neither test below can succeed, and the gather of
                // the match is skipped (every point, in an alignment's first iterations)
                bool may_hold = false;
                if (prior >= 0) {
                    const NNCache m0 = cache[slot];
                    const float ax = c.qx - m0.qx, ay = c.qy - m0.qy, az = c.qz - m0.qz;
                    may_hold = sqrtf(fmaf(az, az, fmaf(ay, ay, ax * ax))) * 1.000001f + pp.pos_err <
                               sqrtf(fabsf(m0.sec)) * 0.999999f;  // NaN (no cache) compares false
                }
                if (may_hold) {
#else
                if (prior >= 0) {
#endif
                    const NNCache m = cache[slot];
                    const float4 t = __ldg(G.hi + prior);
                    const float dx = c.qx - t.x, dy = c.qy - t.y, dz = c.qz - t.z;
                    const float d1 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    const float mx = c.qx - m.qx, my = c.qy - m.qy, mz = c.qz - m.qz;
                    const float moved = sqrtf(fmaf(mz, mz, fmaf(my, my, mx * mx)));
                    // |q - b| (upper bound) + move (upper bound) < distance to anything else then (lower bound);
                    // NaN (no cache) compares false
                    const float lhs = sqrtf(d1 + band(G.p, d1)) * 1.000001f + moved * 1.000001f + pp.pos_err;
                    if (lhs < sqrtf(m.sec) * 0.999999f) {
                        hard = false;
                        // still the nearest, but it may have left the radius: then there is no correspondence
                        // (and nothing to remember: corr_s doubles as the correspondence list)
                        if (!ctx.row_of(prior, vs, pp.r2, x)) corr_s[slot] = -1;
                    } else if (m.sec < 0.0f) {
                        // a two-candidate entry (winner and runner-up too close to tell apart for long): if both
                        // are still nearer than anything else can be, the winner between them is taken in double
                        const int other = second_s[slot];
                        if (other >= 0) {
                            const float4 u = __ldg(G.hi + other);
                            const float ex = c.qx - u.x, ey = c.qy - u.y, ez = c.qz - u.z;
                            const float dm = fmaxf(d1, fmaf(ez, ez, fmaf(ey, ey, ex * ex)));
                            const float lhs2 = sqrtf(dm + band(G.p, dm)) * 1.000001f + moved * 1.000001f + pp.pos_err;
                            if (lhs2 < sqrtf(-m.sec) * 0.999999f) {
                                hard = false;
                                const double da = l2_exact(vs[0], vs[1], vs[2], G.xyz + kPtStride * (int64_t)prior);
                                const double db = l2_exact(vs[0], vs[1], vs[2], G.xyz + kPtStride * (int64_t)other);
                                const bool first = da < db || (da == db && __ldg(G.orig + prior) < __ldg(G.orig + other));
                                if (!first) { corr_s[slot] = other; second_s[slot] = prior; }
                                if (!ctx.row_of(first ? prior : other, vs, pp.r2, x)) corr_s[slot] = -1;
                            }
                        }
                    }
                }
#endif
            }
        }
#ifdef VB_STATS
        if (valid) {
            const int pr = corr_s[task.corr_begin + local];
            VB_STAT(12, hard);
            VB_STAT(13, hard && pr < 0);
            VB_STAT(14, hard && pr >= 0 && !(cache[task.corr_begin + local].sec >= 0.0f));
        }
#endif
        // (Prefetching the next batch's streams and the match's rows ahead of the arithmetic was measured
        // slower — 0.089 vs 0.084 ms for a settled pass: at 48 warps per SM the gathers already overlap.)
        const unsigned hm = __ballot_sync(0xffffffffu, hard);
        if (hard) my_hard[nhard + __popc(hm & ((1u << lane) - 1u))] = (unsigned char)(k * 32 + lane);
        nhard += __popc(hm);
        if (hm != 0xffffffffu) ctx.accumulate(rows_sh[warp], x);  // warp-uniform; nothing to add when every lane is hard
    }
    if (lane == 0) hard_cnt[blockIdx.x * kPassWarps + warp] = nhard;
    // one partial row per block: the warps' Gram matrices summed in warp order (rows_sh is free again)
    const double2 mine = ctx.lane_pair();
    __shared__ int listed[kPassWarps];
    __syncthreads();
    double2 *red = reinterpret_cast<double2 *>(&rows_sh[0][0]);
    red[warp * 32 + lane] = mine;
    if (lane == 0) listed[warp] = nhard;
    __syncthreads();
    double2 *out = reinterpret_cast<double2 *>(partials + (int64_t)blockIdx.x * kRowsPerBlock * kPart);
    if (warp == 0) {
        double2 t = red[lane];
#pragma unroll
        for (int w = 1; w < kPassWarps; w++) { t.x += red[w * 32 + lane].x; t.y += red[w * 32 + lane].y; }
        out[lane] = t;
    }
    // nothing listed: part B's warps for this block exit on their first load, so their rows are zeroed here
    int total = 0;
#pragma unroll
    for (int w = 0; w < kPassWarps; w++) total += listed[w];
    if (total == 0) out[(kPart / 2) + threadIdx.x] = make_double2(0.0, 0.0);  // 4 rows x 32 double2 = 128 threads
#ifdef VB_PB_WORKLIST
    else if (threadIdx.x == 0) pp.work[atomicAdd(pp.work_ctr + 2 * pp.parity, 1)] = blockIdx.x;
#endif
}

// entry t of the 32 estimator slots from a summed 8x8 Gram matrix D.  Slot layout: point-to-plane 0..20 JTJ
// upper triangle row-major, 21..26 JTr; point-to-point 0..2 sum s', 3..5 sum d', 6..14 sum d' s'^T,
// 15 sum |s'|^2; both 30 = sum d2, 31 = count.
__device__ __forceinline__ double slot_from_gram(const double *D, bool plane, int t) {
    double v = 0.0;
    if (t == kSlotD2) v = D[63];
    else if (t == kSlotCount) v = D[62];
    else if (plane) {
        if (t < 21) {
            int a = 0, rem = t;
            while (rem >= 6 - a) { rem -= 6 - a; ++a; }  // slot t = (a, b) of the upper triangle, b >= a
            v = D[8 * a + a + rem];
        } else if (t < 27) {
            v = D[8 * (t - 21) + 6];
        }
    } else {
        if (t < 6) v = D[8 * t + 6];
        else if (t < 15) v = D[8 * (3 + (t - 6) / 3) + (t - 6) % 3];
        else if (t == 15) v = (D[0] + D[9]) + D[18];
    }
    return v;
}

__device__ void solve_from_totals(const double *tot, double npts, ProbState *st, const SolveParams &sp, int pass_index);

// The iteration tail run by the last part-B warp of a problem (kept out of line: it is rare and register-hungry).
static __device__ __noinline__ void iteration_tail(const ProbDesc pd, const double *partials, int *ctr, double *scratch,
                                                   bool plane, ProbState *st, const SolveParams &sp, int pass_index) {
    const int lane = threadIdx.x & 31;
    __threadfence();
    if (lane == 0) *ctr = 0;  // ready for the next pass
    const double2 *rows = reinterpret_cast<const double2 *>(partials + (int64_t)pd.blk_begin * kRowsPerBlock * kPart) + lane;
    const int nrows = pd.blk_count * kRowsPerBlock;
    double ax = 0.0, ay = 0.0;  // entries 2*lane, 2*lane + 1 of the summed Gram matrix; rows added in index order
    int r = 0;
    for (; r + 8 <= nrows; r += 8) {
        double2 v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = __ldcg(rows + (int64_t)(r + i) * (kPart / 2));  // written by other SMs
#pragma unroll
        for (int i = 0; i < 8; i++) { ax += v[i].x; ay += v[i].y; }
    }
    for (; r < nrows; r++) {
        const double2 v = __ldcg(rows + (int64_t)r * (kPart / 2));
        ax += v.x; ay += v.y;
    }
    __syncwarp();
    double *D = scratch, *tot = scratch + kPart;
    D[2 * lane] = ax; D[2 * lane + 1] = ay;
    __syncwarp();
    tot[lane] = slot_from_gram(D, plane, lane);
    __syncwarp();
    if (lane == 0) solve_from_totals(tot, (double)pd.npts, st, sp, pass_index);
}

// ---- pass, part B: the listed points of a part-A block, 32 searches at a time with every lane busy.  Each
// search refreshes the point's cache entry.  One WARP per CUDA block (grid = 4 x part A's): once an alignment
// settles most lists hold a handful of points, and single-warp blocks let an SM keep ~24 of those short,
// latency-bound batches in flight instead of 6 four-warp blocks with one live warp each.
template <int MODE>
__global__ void __launch_bounds__(32, 4 * VB_PASS_MINBLOCKS) k_pass_b(
    GridDev G, const double *__restrict__ src_xyz, const BlockTask *__restrict__ tasks,
    ProbState *states, double *partials, int *__restrict__ corr_s,
    NNCache *__restrict__ cache, int *__restrict__ second_s, const unsigned char *__restrict__ hard_ids,
    const int *__restrict__ hard_cnt, PassParams pp, const ProbDesc *__restrict__ probs, int *prob_ctr, SolveParams sp, int pass_index, int fuse_solve) {
    const int blk = blockIdx.x / kPassWarps, warp = blockIdx.x % kPassWarps, lane = threadIdx.x;
    // the block's list = its warps' lists of part A, concatenated
    int seg_end[kPassWarps];
    int total = 0;
#pragma unroll
    for (int w = 0; w < kPassWarps; w++) {
        total += hard_cnt[blk * kPassWarps + w];
        seg_end[w] = total;
    }
    // Nothing to search (most blocks once an alignment settles): part A has zeroed this warp's partial row.
    // (hard_cnt of a finished problem is stale; its rows are never read again, so either exit is fine.)
    if (total == 0 && !fuse_solve) return;
    const BlockTask task = tasks[blk];
    const ProbState *st = states + task.prob;
    if (st->done) return;
    __shared__ WarpScratch ws;
    PassCtx<MODE> ctx(G, st->T);
#pragma unroll 1
    for (int h0 = 32 * warp; h0 < total; h0 += 32 * kPassWarps) {
        const bool live = h0 + lane < total;
        double x[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        double vs[3] = {0, 0, 0}, d2 = 0.0;
        QueryCtx c;
        int prior = -1, slot = 0;
        float slack = 0.0f;
        if (live) {
            const int h = h0 + lane;
            int w = 0, base = 0;
#pragma unroll
            for (int i = 0; i < kPassWarps - 1; i++)
                if (h >= seg_end[i]) { w = i + 1; base = seg_end[i]; }
            const int id = hard_ids[((int64_t)blk * kPassWarps + w) * (32 * kPtsPerThread) + (h - base)];
            const int local = (w * kPtsPerThread + (id >> 5)) * 32 + (id & 31);
            slot = task.corr_begin + local;
            ctx.transform(src_xyz + 3 * (int64_t)(task.src_begin + local), vs);
            make_query(G.p, vs[0], vs[1], vs[2], c);  // inside the grid, or it would not be listed
            prior = corr_s[slot];  // last iteration's match: a bound for this search
            // looking further than the answer needs only pays when the point is about to settle: one that
            // has just moved by more than the slack will fail the next test anyway
            const NNCache m = cache[slot];
            const float mx = c.qx - m.qx, my = c.qy - m.qy, mz = c.qz - m.qz;
            if (fmaf(mz, mz, fmaf(my, my, mx * mx)) < 4.0f * pp.slack * pp.slack) slack = pp.slack;  // NaN: no
        }
        float sec = -1.0f;
        int other = -1;
#ifdef VB_SEARCH_COOP_ONLY
        const int bs = nn_search_warp(G, live, c, vs[0], vs[1], vs[2], pp.r2, pp.r2_ub, &d2);
#else
        const int bs = nn_search_hybrid<32>(G, live, c, vs[0], vs[1], vs[2], pp.r2, pp.r2_ub, prior, ws.runs, &d2,
                                            slack, &sec, VB_COOP_WARP_LANES, &other);
#endif
        if (live) {
            corr_s[slot] = bs;  // sorted position; vb200_batch_corr maps it to the caller's index
            NNCache m;
            m.qx = c.qx; m.qy = c.qy; m.qz = c.qz;
            m.sec = bs >= 0 ? sec : -1.0f;
            cache[slot] = m;
            second_s[slot] = bs >= 0 ? other : -1;
            if (bs >= 0) ctx.row_of(bs, vs, pp.r2, x);
        }
        ctx.accumulate(ws.rows, x);
    }
    ctx.write_partial(partials + ((int64_t)blk * kRowsPerBlock + 1 + warp) * kPart);
    if (!fuse_solve) return;
    // ---- iteration tail, fused: the LAST warp of a problem to get here (all of the problem's partial rows are
    // then in memory) sums them in a fixed order and takes the estimator step — no separate solve launch.
    __threadfence();
    int prev = 0;
    if (lane == 0) prev = atomicAdd(prob_ctr + task.prob, 1);
    prev = __shfl_sync(0xffffffffu, prev, 0);
    const ProbDesc pd = probs[task.prob];
    if (prev != pd.blk_count * kPassWarps - 1) return;
    iteration_tail(pd, partials, prob_ctr + task.prob, ws.rows, MODE == 1, states + task.prob, sp, pass_index);
}

#ifdef VB_PB_WORKLIST
// ---- pass, part B over a worklist.  k_pass_b above launches one single-warp block per (part-A block, warp):
// 12 544 blocks on the BASELINE workload, nearly all of which find an empty list once an alignment settles, and
// in the first iterations the block scheduler's fixed order leaves a tail.  Here part A appends the blocks that listed anything to a worklist (one atomic per such block) and part B is a
// resident grid of single-warp blocks, each of which takes (block, warp) items off the list one at a time (an
// atomic per item: dynamic balance in the first iterations, when every block is listed) until none is left.
// Which warp handles an item is arbitrary; what it computes and where it writes (the item's own partial row,
// the points' own slots) is not: results are bit for bit those of k_pass_b.
template <int MODE>
__global__ void __launch_bounds__(32, 4 * VB_PASS_MINBLOCKS) k_pass_b_wl(
    GridDev G, const double *__restrict__ src_xyz, const BlockTask *__restrict__ tasks,
    const ProbState *__restrict__ states, double *partials, int *__restrict__ corr_s,
    NNCache *__restrict__ cache, int *__restrict__ second_s, const unsigned char *__restrict__ hard_ids,
    const int *__restrict__ hard_cnt, PassParams pp) {
    const int lane = threadIdx.x;
    __shared__ WarpScratch ws;
    int *ctr = pp.work_ctr + 2 * pp.parity;
    const int n_items = ctr[0] * kPassWarps;  // final: part A has finished
#pragma unroll 1
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(ctr + 1, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        const int blk = pp.work[item / kPassWarps], warp = item % kPassWarps;
        int seg_end[kPassWarps];
        int total = 0;
#pragma unroll
        for (int w = 0; w < kPassWarps; w++) {
            total += hard_cnt[blk * kPassWarps + w];
            seg_end[w] = total;
        }
        const BlockTask task = tasks[blk];
        PassCtx<MODE> ctx(G, states[task.prob].T);  // (a finished problem's blocks are never listed)
#pragma unroll 1
        for (int h0 = 32 * warp; h0 < total; h0 += 32 * kPassWarps) {
            const bool live = h0 + lane < total;
            double x[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            double vs[3] = {0, 0, 0}, d2 = 0.0;
            QueryCtx c;
            int prior = -1, slot = 0;
            float slack = 0.0f;
            if (live) {
                const int h = h0 + lane;
                int w = 0, base = 0;
#pragma unroll
                for (int i = 0; i < kPassWarps - 1; i++)
                    if (h >= seg_end[i]) { w = i + 1; base = seg_end[i]; }
                const int id = hard_ids[((int64_t)blk * kPassWarps + w) * (32 * kPtsPerThread) + (h - base)];
                const int local = (w * kPtsPerThread + (id >> 5)) * 32 + (id & 31);
                slot = task.corr_begin + local;
                ctx.transform(src_xyz + 3 * (int64_t)(task.src_begin + local), vs);
                make_query(G.p, vs[0], vs[1], vs[2], c);
                prior = corr_s[slot];
                const NNCache m = cache[slot];
                const float mx = c.qx - m.qx, my = c.qy - m.qy, mz = c.qz - m.qz;
                if (fmaf(mz, mz, fmaf(my, my, mx * mx)) < 4.0f * pp.slack * pp.slack) slack = pp.slack;
            }
            float sec = -1.0f;
            int other = -1;
            const int bs = nn_search_hybrid<32>(G, live, c, vs[0], vs[1], vs[2], pp.r2, pp.r2_ub, prior, ws.runs, &d2,
                                                slack, &sec, VB_COOP_WARP_LANES, &other);
            if (live) {
                corr_s[slot] = bs;
                NNCache m;
                m.qx = c.qx; m.qy = c.qy; m.qz = c.qz;
                m.sec = bs >= 0 ? sec : -1.0f;
                cache[slot] = m;
                second_s[slot] = bs >= 0 ? other : -1;
                if (bs >= 0) ctx.row_of(bs, vs, pp.r2, x);
            }
            ctx.accumulate(ws.rows, x);
        }
        ctx.write_partial(partials + ((int64_t)blk * kRowsPerBlock + 1 + warp) * kPart);
        __syncwarp();
    }
}
#endif

// ---- estimator solves from the reduced slots ---------------------------------------------------------
__device__ void update_p2plane(const double *tot, const SolveParams &sp, double *U) {
    mat4_identity(U);
    double JTJ[36], JTr[6];
    int k = 0;
#pragma unroll
    for (int a = 0; a < 6; a++) {
#pragma unroll
        for (int b = a; b < 6; b++) { JTJ[6 * a + b] = tot[k]; JTJ[6 * b + a] = tot[k]; k++; }
    }
#pragma unroll
    for (int a = 0; a < 6; a++) JTr[a] = tot[21 + a];
    if (sp.estimator == VB200_EST_P2PLANE) {
        double x[6];
        if (solve_normal_equations<6>(JTJ, JTr, x)) vec6_to_T(x, U);  // else Identity (TransformationEstimation.cpp:102)
    } else {
        // gravity-constrained: J4 = [(vs x nt).g ; nt] = P J6 with P = [[g^T 0],[0 I]]  =>  A4 = P JTJ P^T
        const double *g = sp.g;
        double A[16], b[4];
        double Mg[6];  // JTJ[:, 0:3] * g
        for (int r = 0; r < 6; r++) Mg[r] = JTJ[6 * r + 0] * g[0] + JTJ[6 * r + 1] * g[1] + JTJ[6 * r + 2] * g[2];
        A[0] = g[0] * Mg[0] + g[1] * Mg[1] + g[2] * Mg[2];
        for (int c = 0; c < 3; c++) { A[1 + c] = Mg[3 + c]; A[4 * (1 + c)] = Mg[3 + c]; }
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) A[4 * (1 + r) + 1 + c] = JTJ[6 * (3 + r) + 3 + c];
        b[0] = g[0] * JTr[0] + g[1] * JTr[1] + g[2] * JTr[2];
        for (int c = 0; c < 3; c++) b[1 + c] = JTr[3 + c];
        double x[4];
        if (solve_normal_equations<4>(A, b, x)) axis_angle_to_T(x[0], g, x + 1, U);
    }
}

__device__ void update_p2p(const double *tot, const double *cref, double *U) {
    mat4_identity(U);
    const double K = tot[kSlotCount];
    if (!(K > 0.0)) return;  // corres.empty() -> Identity (TransformationEstimation.cpp:51)
    const double inv = 1.0 / K;
    double ms[3], md[3], Sigma[9];
    for (int a = 0; a < 3; a++) { ms[a] = tot[a] * inv; md[a] = tot[3 + a] * inv; }
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) Sigma[3 * a + b] = tot[6 + 3 * a + b] * inv - md[a] * ms[b];
    double R[9];
    kabsch_rotation(Sigma, R, nullptr);
    // t = mu_d - R mu_s with mu = cref + mu'
    for (int r = 0; r < 3; r++) {
        double mus[3] = {cref[0] + ms[0], cref[1] + ms[1], cref[2] + ms[2]};
        U[4 * r + 0] = R[3 * r + 0]; U[4 * r + 1] = R[3 * r + 1]; U[4 * r + 2] = R[3 * r + 2];
        U[4 * r + 3] = (cref[r] + md[r]) - (R[3 * r] * mus[0] + R[3 * r + 1] * mus[1] + R[3 * r + 2] * mus[2]);
    }
}

// Fixed-order reduction of one problem's per-warp partials (8x8 Gram matrices, kPart doubles each) into the
// 32 estimator slots tot[kAcc] (shared memory); 256 threads.  Slot layout: point-to-plane 0..20 JTJ upper
// triangle row-major, 21..26 JTr; point-to-point 0..2 sum s', 3..5 sum d', 6..14 sum d' s'^T, 15 sum |s'|^2;
// both 30 = sum d2, 31 = count.
__device__ __forceinline__ void reduce_partials(const ProbDesc &pd, const double *__restrict__ partials,
                                                const int *__restrict__ hard_cnt, bool plane,
                                                double (*sw)[kPart], double *tot) {
    const int e = threadIdx.x & (kPart - 1), grp = threadIdx.x / kPart;  // 4 groups of 64 threads
    double s = 0.0;
    const double *col = partials + (int64_t)pd.blk_begin * kRowsPerBlock * kPart + e;
#ifdef VB_SOLVE_SKIP
    // Part-B rows of blocks whose list was empty are zeros (k_pass_a wrote them): not loading them leaves every
    // sum's bits alone (s + 0.0 == s; s is never -0.0) and, once an alignment settles, four fifths of the rows
    // are such zeros.  Which blocks listed anything is read ONCE per chunk of blocks into shared memory — one
    // coalesced 16-byte load per block — so no row has a dependent global load in front of it (the first attempt
    // looked the count up per row and was slower than loading the zeros).  Rows keep their group and their
    // position in the chain of adds: a chunk is a whole number of 64-row steps.
    constexpr int kFlagChunk = 2048;  // blocks per chunk
    static_assert(kPassWarps == 4 && (kFlagChunk * kRowsPerBlock) % (4 * 16) == 0, "chunk = whole unrolled steps");
    __shared__ unsigned char listed[kFlagChunk];
    for (int c0 = 0; c0 < pd.blk_count; c0 += kFlagChunk) {
        const int nb = min(kFlagChunk, pd.blk_count - c0);
        __syncthreads();  // the previous chunk's flags have been consumed
        for (int j = threadIdx.x; j < nb; j += blockDim.x) {
            const int4 h = reinterpret_cast<const int4 *>(hard_cnt)[pd.blk_begin + c0 + j];
            listed[j] = (h.x | h.y | h.z | h.w) != 0;
        }
        __syncthreads();
        const int r0 = c0 * kRowsPerBlock, r1 = r0 + nb * kRowsPerBlock;
        for (int b = r0 + grp; b < r1; b += 4 * 16) {
            double v[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int r = b + 4 * i;
                const int blk = r / kRowsPerBlock;
                const bool need = r < r1 && (r == blk * kRowsPerBlock || listed[min(blk - c0, kFlagChunk - 1)]);
                v[i] = need ? col[(int64_t)r * kPart] : 0.0;
            }
#pragma unroll
            for (int i = 0; i < 16; i++) s += v[i];
        }
    }
#else
    // rows in index order per group.  (Skipping the part-B rows of blocks with an empty list was measured
    // slower: the count lookup puts a dependent load in front of every row.)
    (void)hard_cnt;
    // kSolveInflight independent loads in flight per thread (the rows sit in L2; the chain of adds keeps its fixed
    // order whatever the batch size, so this knob never changes a bit of the result)
#ifndef VB_SOLVE_INFLIGHT
#define VB_SOLVE_INFLIGHT 16
#endif
    constexpr int kSolveInflight = VB_SOLVE_INFLIGHT;
    const int nrows = pd.blk_count * kRowsPerBlock;
    // (the ragged last batch is predicated, not a loop of dependent loads: adding +0.0 leaves the sum's bits alone)
    for (int b = grp; b < nrows; b += 4 * kSolveInflight) {
        double v[kSolveInflight];
#pragma unroll
        for (int i = 0; i < kSolveInflight; i++) v[i] = b + 4 * i < nrows ? col[(int64_t)(b + 4 * i) * kPart] : 0.0;
#pragma unroll
        for (int i = 0; i < kSolveInflight; i++) s += v[i];
    }
#endif
    sw[grp][e] = s;
    __syncthreads();
    if (threadIdx.x < kPart) sw[0][e] = ((sw[0][e] + sw[1][e]) + sw[2][e]) + sw[3][e];
    __syncthreads();
    if (threadIdx.x < kAcc) tot[threadIdx.x] = slot_from_gram(sw[0], plane, threadIdx.x);
    __syncthreads();
}

// result bookkeeping (Registration.cpp:87-93), convergence test (:179-183) and estimator update (:172-174)
// from the reduced slots; one thread.  npts = source points the totals were accumulated over.
__device__ void solve_from_totals(const double *tot, double npts, ProbState *st, const SolveParams &sp, int pass_index) {
    const double K = tot[kSlotCount];
    double fitness = 0.0, rmse = 0.0;
    if (K > 0.0 && npts > 0.0) {
        fitness = K / npts;
        rmse = sqrt(tot[kSlotD2] / K);
    }
    st->fitness = fitness;
    st->rmse = rmse;
    st->ncorr = (int)K;
    if (pass_index >= 1 && fabs(st->prev_fitness - fitness) < sp.rel_fitness &&
        fabs(st->prev_rmse - rmse) < sp.rel_rmse) {
        st->done = 1;
        if (sp.ndone) atomicAdd(sp.ndone, 1);
        return;
    }
    st->prev_fitness = fitness;
    st->prev_rmse = rmse;
    if (pass_index >= sp.max_iter) {
        st->done = 1;
        if (sp.ndone) atomicAdd(sp.ndone, 1);
        return;
    }
    double U[16], T[16];
    for (int i = 0; i < 16; i++) T[i] = st->T[i];
    if (sp.estimator == VB200_EST_P2P) {
        const double cref[3] = {T[3], T[7], T[11]};
        update_p2p(tot, cref, U);
    } else {
        update_p2plane(tot, sp, U);
    }
    mat4_mul(U, T, T);
    for (int i = 0; i < 16; i++) st->T[i] = T[i];
    st->iters = pass_index + 1;
}

// One block per problem: reduce + solve (the single-GPU iteration tail).
__global__ void __launch_bounds__(256) k_solve(const ProbDesc *__restrict__ probs, ProbState *__restrict__ states,
                                               const double *__restrict__ partials,
                                               const int *__restrict__ hard_cnt, SolveParams sp, int pass_index) {
    const ProbDesc pd = probs[blockIdx.x];
    ProbState *st = states + blockIdx.x;
    if (st->done) return;
    __shared__ double sw[4][kPart];
    __shared__ double tot[kAcc];
    reduce_partials(pd, partials, hard_cnt, sp.estimator != VB200_EST_P2P, sw, tot);
    if (threadIdx.x == 0) solve_from_totals(tot, (double)pd.npts, st, sp, pass_index);
}

// The same tail split in two for multi-GPU alignment of ONE cloud sharded over ranks (ICPRefinement's global
// transform, src/evaluation.cpp:244-274): k_reduce leaves each problem's 32 totals in a device buffer the
// caller all-reduces across GPUs, k_solve_totals finishes the iteration from the combined totals.  Every rank
// sees identical totals, hence applies the identical update: no broadcast of T is needed.
__global__ void __launch_bounds__(256) k_reduce(const ProbDesc *__restrict__ probs, const ProbState *__restrict__ states,
                                                const double *__restrict__ partials,
                                                const int *__restrict__ hard_cnt, bool plane,
                                                double *__restrict__ totals) {
    const ProbDesc pd = probs[blockIdx.x];
    __shared__ double sw[4][kPart];
    __shared__ double tot[kAcc];
    if (states[blockIdx.x].done) {  // finished problems contribute their last totals unchanged
        return;
    }
    reduce_partials(pd, partials, hard_cnt, plane, sw, tot);
    if (threadIdx.x < kAcc) totals[(int64_t)blockIdx.x * kAcc + threadIdx.x] = tot[threadIdx.x];
}

__global__ void __launch_bounds__(32) k_solve_totals(int P, ProbState *__restrict__ states,
                                                     const double *__restrict__ totals,
                                                     const double *__restrict__ npts_global, SolveParams sp,
                                                     int pass_index) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P || states[p].done) return;
    double tot[kAcc];
    for (int i = 0; i < kAcc; i++) tot[i] = totals[(int64_t)p * kAcc + i];
    solve_from_totals(tot, npts_global[p], states + p, sp, pass_index);
}

// ---- source-cloud bucket sort (spatial coherence for the search; deterministic order) ----------------
__device__ __forceinline__ int find_cloud(const int *__restrict__ off, int ncloud, int i) {
    int lo = 0, hi = ncloud;  // off[lo] <= i < off[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_src_keys(const double *__restrict__ xyz, int n,
                                                  const int *__restrict__ cloud_off, int ncloud,
                                                  double inv_half_cell, int *__restrict__ key,
                                                  int *__restrict__ counts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = find_cloud(cloud_off, ncloud, i);
    // 5 bits per axis of the half-coarse-cell index in the cloud's own frame: neighbours in space share
    // or neighbour buckets whatever rigid transform is applied later
    int kx = (int)floor(xyz[3 * (int64_t)i] * inv_half_cell) & 31;
    int ky = (int)floor(xyz[3 * (int64_t)i + 1] * inv_half_cell) & 31;
    int kz = (int)floor(xyz[3 * (int64_t)i + 2] * inv_half_cell) & 31;
    int kk = c * kBuckets + ((kz * 32 + ky) * 32 + kx);
    key[i] = kk;
    atomicAdd(counts + kk, 1);
}

// sidx entries carry the point's fine sub-cell (1 bit per axis of the quarter-coarse index) above the point
// index so the per-bucket sort orders a bucket by (sub-cell, index): fine-cell-sized spatial coherence
constexpr int kSubShift = 28;

__global__ void __launch_bounds__(256) k_src_scatter(const double *__restrict__ xyz, const int *__restrict__ key,
                                                     int n, double inv_fine, const int *__restrict__ start,
                                                     int *__restrict__ cursor, int *__restrict__ sidx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int kk = key[i];
    int sx = (int)floor(xyz[3 * (int64_t)i] * inv_fine) & 1;
    int sy = (int)floor(xyz[3 * (int64_t)i + 1] * inv_fine) & 1;
    int sz = (int)floor(xyz[3 * (int64_t)i + 2] * inv_fine) & 1;
    sidx[start[kk] + atomicAdd(cursor + kk, 1)] = ((sz * 4 + sy * 2 + sx) << kSubShift) | i;
}

// Most of the ncloud x 32768 buckets are empty (a warp per bucket spent 0.34 ms on finding that out): one thread
// per bucket lists the ones holding more than one point (any order: every bucket is sorted on its own) ...
__global__ void __launch_bounds__(256) k_src_list_buckets(int nbuckets, const int *__restrict__ start,
                                                          int *__restrict__ list, int *__restrict__ nlist) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const bool need = b < nbuckets && start[b + 1] - start[b] > 1;
    const unsigned m = __ballot_sync(0xffffffffu, need);
    if (m == 0u) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(nlist, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (need) list[base + __popc(m & ((1u << lane) - 1u))] = b;
}

// ... and a fixed grid of warps sorts the listed buckets
__global__ void __launch_bounds__(256) k_src_sort_buckets(const int *__restrict__ list, const int *__restrict__ nlist,
                                                          const int *__restrict__ start, int *__restrict__ sidx) {
    const int n = *nlist;
    const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    for (int i = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); i < n; i += nwarps) {
        const int b = list[i];
        const int s0 = start[b], s1 = start[b + 1];
        warp_cell_sort<int, 8>(sidx + s0, s1 - s0);
    }
}

__global__ void __launch_bounds__(256) k_src_gather(const double *__restrict__ in, int n,
                                                    const int *__restrict__ sidx,
                                                    const int *__restrict__ cloud_off, int ncloud,
                                                    double *__restrict__ out, int *__restrict__ orig) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int i = sidx[s] & ((1 << kSubShift) - 1);
    out[3 * (int64_t)s] = in[3 * (int64_t)i];
    out[3 * (int64_t)s + 1] = in[3 * (int64_t)i + 1];
    out[3 * (int64_t)s + 2] = in[3 * (int64_t)i + 2];
    orig[s] = i - cloud_off[find_cloud(cloud_off, ncloud, i)];
}

// correspondences are kept as sorted scene positions (they double as the next pass's search bound); this maps
// one problem's column to the caller's target indices
__global__ void __launch_bounds__(256) k_corr_orig(const int *__restrict__ corr_s, int n, const int *__restrict__ orig,
                                                   int *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = corr_s[i];
    out[i] = s >= 0 ? orig[s] : -1;
}

// ---- estimator plug-in kernel: reductions over an explicit correspondence list ------------------------
template <int MODE>
__global__ void __launch_bounds__(kPassTpb) k_estimate(const double *__restrict__ vs_in,
                                                       const double *__restrict__ vt_in,
                                                       const double *__restrict__ nt_in, int64_t K,
                                                       double *__restrict__ partials) {
    __shared__ double swarp[kPassTpb / 32][kAcc];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double cref[3] = {vs_in[0], vs_in[1], vs_in[2]};
    double acc = 0.0;
#pragma unroll 1
    for (int k = 0; k < kPtsPerThread; k++) {
        int64_t i = (int64_t)blockIdx.x * kChunk + k * kPassTpb + threadIdx.x;
        bool valid = i < K;
        double vs[3] = {0, 0, 0}, vt[3] = {0, 0, 0}, nt[3] = {0, 0, 0};
        if (valid) {
            for (int a = 0; a < 3; a++) {
                vs[a] = vs_in[3 * i + a];
                vt[a] = vt_in[3 * i + a];
                if (MODE == 1) nt[a] = nt_in[3 * i + a];
            }
        }
        double v[kAcc];
        contributions<MODE>(valid, 0.0, vs, vt, nt, cref, v);
        acc += warp_reduce_slots(v);
    }
    swarp[warp][lane] = acc;
    __syncthreads();
    if (threadIdx.x < kAcc) {
        double s = swarp[0][threadIdx.x];
        for (int w = 1; w < kPassTpb / 32; w++) s += swarp[w][threadIdx.x];
        partials[(int64_t)blockIdx.x * kAcc + threadIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256) k_estimate_solve(const double *__restrict__ partials, int nblk,
                                                        const double *__restrict__ vs_in, SolveParams sp,
                                                        double *__restrict__ out_T) {
    __shared__ double sw[8][kAcc];
    __shared__ double tot[kAcc];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double s = 0.0;
    for (int b = warp; b < nblk; b += 8) s += partials[(int64_t)b * kAcc + lane];
    sw[warp][lane] = s;
    __syncthreads();
    if (threadIdx.x < kAcc) {
        double t = sw[0][threadIdx.x];
        for (int w = 1; w < 8; w++) t += sw[w][threadIdx.x];
        tot[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    double U[16];
    if (sp.estimator == VB200_EST_P2P) {
        const double cref[3] = {vs_in[0], vs_in[1], vs_in[2]};
        update_p2p(tot, cref, U);
    } else {
        update_p2plane(tot, sp, U);
    }
    for (int i = 0; i < 16; i++) out_T[i] = U[i];
}

}  // namespace

// ======================================================================================================
struct Batch {
    Scene *scene = nullptr;
    cudaStream_t stream = nullptr;  // the scene's stream, or its second one (vb200_icp_run pipelines two halves)
    int ncloud = 0;
    int64_t npts = 0;
    std::vector<int> cloud_off;     // host copy, ncloud+1
    bool has_normals = false;
    double *d_src = nullptr;        // sorted source points, 3*npts
    int *d_src_orig = nullptr;      // sorted position -> original index local to its cloud
    int *d_cloud_off = nullptr;
    // problems
    int P = 0;
    bool all_nonempty = false;               // every problem has source points (the fused iteration tail needs one)
    std::vector<ProbDesc> probs;
    int nblk = 0;
    int64_t ncorr_slots = 0;
    ProbDesc *d_probs = nullptr;
    ProbState *d_states = nullptr;
    BlockTask *d_tasks = nullptr;
    double *d_partials = nullptr;
    int *d_corr = nullptr;
    NNCache *d_cache = nullptr;              // per (problem, point): what its last search proved (k_pass_a/b)
    int *d_second = nullptr;                 // ... and the runner-up kept with a two-candidate entry
    unsigned char *d_hard_ids = nullptr;     // per block and warp: the points part A left for part B
    int *d_hard_cnt = nullptr;
    int *d_prob_ctr = nullptr;               // per problem: part-B warps finished in the current pass
    int *d_ndone = nullptr;                  // problems finished since set_problems
#ifdef VB_PB_WORKLIST
    int *d_work = nullptr;                   // part-A blocks with a non-empty list (k_pass_b_wl)
    int *d_work_ctr = nullptr;               // 2 x {listed, handed out}
    int pass_parity = 0;
#endif
    int64_t launches = 0;
    int iter_base = 0;                       // running pass index for vb200_batch_iterate
    double *d_totals = nullptr;              // P x kAcc, library-owned unless the caller supplied a buffer
    double *d_totals_ext = nullptr;
    double *d_npts_global = nullptr;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};  // pass start / pass end = solve start / solve end
    bool ev_valid = false;
};

static void batch_free_problems(Batch *b) {
    cudaStream_t st = b->stream;
    void *ptrs[10] = {b->d_probs, b->d_states, b->d_tasks, b->d_partials, b->d_corr, b->d_totals, b->d_npts_global,
                      b->d_cache, b->d_hard_ids, b->d_hard_cnt};
    if (b->d_prob_ctr) cudaFreeAsync(b->d_prob_ctr, st);
    if (b->d_second) cudaFreeAsync(b->d_second, st);
    b->d_second = nullptr;
    if (b->d_ndone) cudaFreeAsync(b->d_ndone, st);
    b->d_ndone = nullptr;
#ifdef VB_PB_WORKLIST
    if (b->d_work) cudaFreeAsync(b->d_work, st);
    if (b->d_work_ctr) cudaFreeAsync(b->d_work_ctr, st);
    b->d_work = nullptr; b->d_work_ctr = nullptr;
#endif
    b->d_totals = nullptr; b->d_npts_global = nullptr; b->d_cache = nullptr;
    b->d_hard_ids = nullptr; b->d_hard_cnt = nullptr; b->d_prob_ctr = nullptr;
    for (void *q : ptrs)
        if (q) cudaFreeAsync(q, st);
    b->d_probs = nullptr; b->d_states = nullptr; b->d_tasks = nullptr; b->d_partials = nullptr; b->d_corr = nullptr;
    b->P = 0; b->nblk = 0; b->probs.clear();
}

static void batch_free(Batch *b) {
    if (!b) return;
    batch_free_problems(b);
    for (int i = 0; i < 3; i++)
        if (b->ev[i]) cudaEventDestroy(b->ev[i]);
    cudaStream_t st = b->stream;
    void *ptrs[3] = {b->d_src, b->d_src_orig, b->d_cloud_off};
    for (void *q : ptrs)
        if (q) cudaFreeAsync(q, st);
    delete b;
}

static int batch_upload(Batch *b, const double *src_xyz, const int64_t *off, int ncloud) {
    Scene *sc = b->scene;
    cudaStream_t st = b->stream;
    if (ncloud > 32768) return VB200_ERR_INVALID;  // bucket table = ncloud x 32768 counters (int indexed)
    b->ncloud = ncloud;
    b->cloud_off.resize((size_t)ncloud + 1);
    for (int c = 0; c <= ncloud; c++) {
        int64_t o = off[c] - off[0];
        if (o < 0 || o > 0x7fffffff || (c > 0 && off[c] < off[c - 1])) return VB200_ERR_INVALID;
        b->cloud_off[c] = (int)o;
    }
    const int n = b->cloud_off[ncloud];
    if (n >= (1 << kSubShift)) return VB200_ERR_INVALID;  // 2^28 source points per batch
    b->npts = n;
    VB_CUDA(cudaMallocAsync((void **)&b->d_cloud_off, sizeof(int) * ((size_t)ncloud + 1), st));
    VB_CUDA(cudaMemcpyAsync(b->d_cloud_off, b->cloud_off.data(), sizeof(int) * ((size_t)ncloud + 1),
                            cudaMemcpyHostToDevice, st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_src, sizeof(double) * 3 * (size_t)std::max(n, 1), st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_src_orig, sizeof(int) * (size_t)std::max(n, 1), st));
    if (n == 0) return VB200_OK;
    DevBuf<double> d_in(st);
    DevBuf<int> d_key(st), d_counts(st), d_start(st), d_sidx(st);
    const size_t nb = (size_t)ncloud * kBuckets + 1;
    VB_CUDA(d_in.alloc(3 * (size_t)n));
    VB_CUDA(d_key.alloc((size_t)n));
    VB_CUDA(d_sidx.alloc((size_t)n));
    VB_CUDA(d_counts.alloc(nb));
    VB_CUDA(d_start.alloc(nb));
    VB_CUDA(cudaMemcpyAsync(d_in.p, src_xyz + 3 * off[0], sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, st));
    VB_CUDA(cudaMemsetAsync(d_counts.p, 0, sizeof(int) * nb, st));
    const double inv_half = 2.0 / sc->grid.p.cell;
    k_src_keys<<<div_up(n, 256), 256, 0, st>>>(d_in.p, n, b->d_cloud_off, ncloud, inv_half, d_key.p, d_counts.p);
    VB_TRY(exclusive_scan_i32(d_counts.p, d_start.p, (int64_t)nb, nullptr, st));
    VB_CUDA(cudaMemsetAsync(d_counts.p, 0, sizeof(int) * nb, st));
    k_src_scatter<<<div_up(n, 256), 256, 0, st>>>(d_in.p, d_key.p, n, 4.0 / sc->grid.p.cell, d_start.p, d_counts.p, d_sidx.p);
    // the key array is free once the scatter has run: it becomes the list of buckets to sort (at most n / 2)
    VB_CUDA(cudaMemsetAsync(d_counts.p, 0, sizeof(int), st));
    k_src_list_buckets<<<div_up((int64_t)nb - 1, 256), 256, 0, st>>>((int)nb - 1, d_start.p, d_key.p, d_counts.p);
    k_src_sort_buckets<<<kNumSMsB200 * 8, 256, 0, st>>>(d_key.p, d_counts.p, d_start.p, d_sidx.p);
    k_src_gather<<<div_up(n, 256), 256, 0, st>>>(d_in.p, n, d_sidx.p, b->d_cloud_off, ncloud, b->d_src, b->d_src_orig);
    VB_CUDA(cudaGetLastError());
    b->launches += 5 + 3;
    VB_CUDA(cudaStreamSynchronize(st));  // temporaries are released on return
    return VB200_OK;
}

static int batch_set_problems(Batch *b, const int32_t *cloud_ids, const double *init_T, int P) {
    cudaStream_t st = b->stream;
    batch_free_problems(b);
    b->P = P;
    b->iter_base = 0;
    b->probs.resize((size_t)P);
    std::vector<BlockTask> tasks;
    std::vector<ProbState> states((size_t)P);
    int64_t corr = 0;
    for (int p = 0; p < P; p++) {
        int c = cloud_ids ? cloud_ids[p] : p;
        if (c < 0 || c >= b->ncloud) return VB200_ERR_INVALID;
        ProbDesc &pd = b->probs[p];
        pd.cloud = c;
        pd.npts = b->cloud_off[c + 1] - b->cloud_off[c];
        pd.src_begin = b->cloud_off[c];
        pd.corr_begin = (int)corr;
        pd.blk_begin = (int)tasks.size();
        for (int s = 0; s < pd.npts; s += kChunk) {
            BlockTask t;
            t.prob = p;
            t.src_begin = pd.src_begin + s;
            t.count = std::min(kChunk, pd.npts - s);
            t.corr_begin = pd.corr_begin + s;
            tasks.push_back(t);
        }
        pd.blk_count = (int)tasks.size() - pd.blk_begin;
        corr += pd.npts;
        if (corr > 0x7fffffff) return VB200_ERR_INVALID;
        ProbState &s = states[p];
        memset(&s, 0, sizeof(s));
        for (int i = 0; i < 16; i++) s.T[i] = init_T[16 * (size_t)p + i];
    }
    b->nblk = (int)tasks.size();
    b->ncorr_slots = corr;
    // Fusing the iteration tail into part B (its last warp per problem reduces + solves) was measured SLOWER
    // than a separate k_solve launch: one warp summing ~800 partial rows is latency-bound (~50 us vs ~25 us).
    // The code path stays for experiments (-DVB_FUSE_SOLVE).
#ifdef VB_FUSE_SOLVE
    b->all_nonempty = P > 0;
    for (int p = 0; p < P; p++) b->all_nonempty = b->all_nonempty && b->probs[p].npts > 0;
#else
    b->all_nonempty = false;
#endif
    VB_CUDA(cudaMallocAsync((void **)&b->d_probs, sizeof(ProbDesc) * (size_t)std::max(P, 1), st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_states, sizeof(ProbState) * (size_t)std::max(P, 1), st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_tasks, sizeof(BlockTask) * (size_t)std::max(b->nblk, 1), st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_partials, sizeof(double) * kPart * kRowsPerBlock * (size_t)std::max(b->nblk, 1), st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_corr, sizeof(int) * (size_t)std::max<int64_t>(corr, 1), st));
    if (P) {
        VB_CUDA(cudaMemcpyAsync(b->d_probs, b->probs.data(), sizeof(ProbDesc) * (size_t)P, cudaMemcpyHostToDevice, st));
        VB_CUDA(cudaMemcpyAsync(b->d_states, states.data(), sizeof(ProbState) * (size_t)P, cudaMemcpyHostToDevice, st));
    }
    if (b->nblk)
        VB_CUDA(cudaMemcpyAsync(b->d_tasks, tasks.data(), sizeof(BlockTask) * (size_t)b->nblk, cudaMemcpyHostToDevice, st));
    VB_CUDA(cudaMemsetAsync(b->d_corr, 0xff, sizeof(int) * (size_t)std::max<int64_t>(corr, 1), st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_cache, sizeof(NNCache) * (size_t)std::max<int64_t>(corr, 1), st));
    VB_CUDA(cudaMemsetAsync(b->d_cache, 0xff, sizeof(NNCache) * (size_t)std::max<int64_t>(corr, 1), st));  // NaN
    VB_CUDA(cudaMallocAsync((void **)&b->d_second, sizeof(int) * (size_t)std::max<int64_t>(corr, 1), st));
    VB_CUDA(cudaMemsetAsync(b->d_second, 0xff, sizeof(int) * (size_t)std::max<int64_t>(corr, 1), st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_hard_ids, (size_t)kChunk * (size_t)std::max(b->nblk, 1), st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_hard_cnt, sizeof(int) * kPassWarps * (size_t)std::max(b->nblk, 1), st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_prob_ctr, sizeof(int) * (size_t)std::max(P, 1), st));
    VB_CUDA(cudaMemsetAsync(b->d_prob_ctr, 0, sizeof(int) * (size_t)std::max(P, 1), st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_ndone, sizeof(int), st));
    VB_CUDA(cudaMemsetAsync(b->d_ndone, 0, sizeof(int), st));
#ifdef VB_PB_WORKLIST
    VB_CUDA(cudaMallocAsync((void **)&b->d_work, sizeof(int) * (size_t)std::max(b->nblk, 1), st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_work_ctr, sizeof(int) * 4, st));
    VB_CUDA(cudaMemsetAsync(b->d_work_ctr, 0, sizeof(int) * 4, st));
    b->pass_parity = 0;
#endif
    VB_CUDA(cudaStreamSynchronize(st));  // host vectors go out of scope
    return VB200_OK;
}


// one correspondence pass = part A (every point, streaming) + part B (the points that need a search).  With
// `sp` given, part B also finishes the iteration (reduction + estimator step by each problem's last warp).
static void launch_pass(Batch *b, bool plane, const PassParams &pp_in, const SolveParams *sp = nullptr, int pass_index = 0) {
    Scene *sc = b->scene;
    cudaStream_t st = b->stream;
#ifdef VB_PB_WORKLIST
    PassParams pp = pp_in;
    pp.work = b->d_work;
    pp.work_ctr = b->d_work_ctr;
    pp.parity = b->pass_parity;
    b->pass_parity ^= 1;
    (void)sp; (void)pass_index;
    const int nwarps = std::min(b->nblk * kPassWarps, kNumSMsB200 * 4 * VB_PASS_MINBLOCKS);
    if (plane) {
        k_pass_a<1><<<b->nblk, kPassTpb, 0, st>>>(sc->grid, b->d_src, b->d_tasks, b->d_states, b->d_partials, b->d_corr,
                                                  b->d_cache, b->d_second, b->d_hard_ids, b->d_hard_cnt, pp);
        k_pass_b_wl<1><<<nwarps, 32, 0, st>>>(sc->grid, b->d_src, b->d_tasks, b->d_states, b->d_partials, b->d_corr,
                                              b->d_cache, b->d_second, b->d_hard_ids, b->d_hard_cnt, pp);
    } else {
        k_pass_a<0><<<b->nblk, kPassTpb, 0, st>>>(sc->grid, b->d_src, b->d_tasks, b->d_states, b->d_partials, b->d_corr,
                                                  b->d_cache, b->d_second, b->d_hard_ids, b->d_hard_cnt, pp);
        k_pass_b_wl<0><<<nwarps, 32, 0, st>>>(sc->grid, b->d_src, b->d_tasks, b->d_states, b->d_partials, b->d_corr,
                                              b->d_cache, b->d_second, b->d_hard_ids, b->d_hard_cnt, pp);
    }
    b->launches += 2;
    return;
#else
    const PassParams &pp = pp_in;
#endif
    SolveParams s0;
    memset(&s0, 0, sizeof(s0));
    const SolveParams &s = sp ? *sp : s0;
    const int fuse = sp ? 1 : 0;
    if (plane) {
        k_pass_a<1><<<b->nblk, kPassTpb, 0, st>>>(sc->grid, b->d_src, b->d_tasks, b->d_states, b->d_partials, b->d_corr,
                                                  b->d_cache, b->d_second, b->d_hard_ids, b->d_hard_cnt, pp);
        k_pass_b<1><<<b->nblk * kPassWarps, 32, 0, st>>>(sc->grid, b->d_src, b->d_tasks, b->d_states, b->d_partials,
                                                         b->d_corr, b->d_cache, b->d_second, b->d_hard_ids, b->d_hard_cnt, pp,
                                                         b->d_probs, b->d_prob_ctr, s, pass_index, fuse);
    } else {
        k_pass_a<0><<<b->nblk, kPassTpb, 0, st>>>(sc->grid, b->d_src, b->d_tasks, b->d_states, b->d_partials, b->d_corr,
                                                  b->d_cache, b->d_second, b->d_hard_ids, b->d_hard_cnt, pp);
        k_pass_b<0><<<b->nblk * kPassWarps, 32, 0, st>>>(sc->grid, b->d_src, b->d_tasks, b->d_states, b->d_partials,
                                                         b->d_corr, b->d_cache, b->d_second, b->d_hard_ids, b->d_hard_cnt, pp,
                                                         b->d_probs, b->d_prob_ctr, s, pass_index, fuse);
    }
    b->launches += 2;
}

// Iterations [it_begin, it_end] of the loop (the whole loop is 0..max_iter): vb200_icp_run enqueues the first few
// of one half of its clouds, uploads the other half meanwhile, then comes back for the rest.
static int batch_run(Batch *b, int estimator, const double *gravity, double max_dist, double rel_fitness,
                     double rel_rmse, int max_iter, int it_begin = 0, int it_end = 0x7fffffff) {
    Scene *sc = b->scene;
    cudaStream_t st = b->stream;
    if (estimator < VB200_EST_P2P || estimator > VB200_EST_P2PLANE_GRAVITY || max_iter < 0) return VB200_ERR_INVALID;
    if (!(max_dist > 0.0)) return VB200_ERR_DISTANCE;
    if (max_dist > sc->grid.p.cell * (1.0 + 1e-12)) return VB200_ERR_INVALID;
    const bool plane = estimator != VB200_EST_P2P;
    if (plane && (!b->has_normals || !sc->has_normals)) return VB200_ERR_NORMALS;
    if (b->P == 0) return VB200_OK;
    const PassParams pp = make_pass_params(sc->grid.p, max_dist);
    SolveParams sp;
    sp.ndone = nullptr;
    sp.rel_fitness = rel_fitness;
    sp.rel_rmse = rel_rmse;
    sp.max_iter = max_iter;
    sp.estimator = estimator;
    sp.g[0] = 0.0; sp.g[1] = 1.0; sp.g[2] = 0.0;  // VISMA's gravity convention: +Y (src/annotation.cpp:43,84)
    if (estimator == VB200_EST_P2PLANE_GRAVITY && gravity) {
        double l = sqrt(gravity[0] * gravity[0] + gravity[1] * gravity[1] + gravity[2] * gravity[2]);
        if (!(l > 0.0)) return VB200_ERR_INVALID;
        for (int a = 0; a < 3; a++) sp.g[a] = gravity[a] / l;
    }
    // With a live convergence test most runs finish long before max_iter (5-7 iterations on the BASELINE
    // workload): every fourth iteration the host looks at the finished-problems counter and stops enqueueing
    // passes that would only find every problem done.
    const bool can_stop = rel_fitness > 0.0 && rel_rmse > 0.0;
    if (can_stop) sp.ndone = b->d_ndone;
    for (int it = it_begin; it <= std::min(max_iter, it_end); it++) {
        if (can_stop && it >= 4 && (it & 3) == 0) {
            int ndone = 0;
            VB_CUDA(cudaMemcpyAsync(&ndone, b->d_ndone, sizeof(int), cudaMemcpyDeviceToHost, st));
            VB_CUDA(cudaStreamSynchronize(st));
            if (ndone >= b->P) break;
        }
        if (b->all_nonempty) {
            launch_pass(b, plane, pp, &sp, it);
        } else {
            // a problem without source points has no part-B warp to finish it: separate solve launch
            if (b->nblk) launch_pass(b, plane, pp);
            k_solve<<<b->P, 256, 0, st>>>(b->d_probs, b->d_states, b->d_partials, b->d_hard_cnt, sp, it);
            b->launches++;
        }
    }
    VB_CUDA(cudaGetLastError());
    return VB200_OK;
}

static int make_params(Batch *b, int estimator, const double *gravity, double max_dist, PassParams *pp,
                       SolveParams *sp) {
    Scene *sc = b->scene;
    if (estimator < VB200_EST_P2P || estimator > VB200_EST_P2PLANE_GRAVITY) return VB200_ERR_INVALID;
    if (!(max_dist > 0.0)) return VB200_ERR_DISTANCE;
    if (max_dist > sc->grid.p.cell * (1.0 + 1e-12)) return VB200_ERR_INVALID;
    if (estimator != VB200_EST_P2P && (!b->has_normals || !sc->has_normals)) return VB200_ERR_NORMALS;
    *pp = make_pass_params(sc->grid.p, max_dist);
    sp->estimator = estimator;
    sp->g[0] = 0.0; sp->g[1] = 1.0; sp->g[2] = 0.0;
    if (estimator == VB200_EST_P2PLANE_GRAVITY && gravity) {
        double l = sqrt(gravity[0] * gravity[0] + gravity[1] * gravity[1] + gravity[2] * gravity[2]);
        if (!(l > 0.0)) return VB200_ERR_INVALID;
        for (int a = 0; a < 3; a++) sp->g[a] = gravity[a] / l;
    }
    return VB200_OK;
}

// first half of an iteration: correspondence pass + per-problem totals left on the device
static int batch_pass(Batch *b, int estimator, double max_dist) {
    cudaStream_t st = b->stream;
    PassParams pp;
    SolveParams sp;
    sp.ndone = nullptr;
    VB_TRY(make_params(b, estimator, nullptr, max_dist, &pp, &sp));
    if (b->P == 0) return VB200_OK;
    if (!b->d_totals && !b->d_totals_ext) {
        VB_CUDA(cudaMallocAsync((void **)&b->d_totals, sizeof(double) * kAcc * (size_t)b->P, st));
        VB_CUDA(cudaMemsetAsync(b->d_totals, 0, sizeof(double) * kAcc * (size_t)b->P, st));
    }
    double *totals = b->d_totals_ext ? b->d_totals_ext : b->d_totals;
    if (b->nblk) {
        launch_pass(b, estimator != VB200_EST_P2P, pp);
    }
    k_reduce<<<b->P, 256, 0, st>>>(b->d_probs, b->d_states, b->d_partials, b->d_hard_cnt, estimator != VB200_EST_P2P, totals);
    b->launches++;
    VB_CUDA(cudaGetLastError());
    return VB200_OK;
}

// second half: finish the iteration from the (all-reduced) totals
static int batch_solve(Batch *b, int estimator, const double *gravity, double max_dist, double rel_fitness,
                       double rel_rmse, int max_iter, int pass_index, const int64_t *npts_global) {
    cudaStream_t st = b->stream;
    PassParams pp;
    SolveParams sp;
    sp.ndone = nullptr;
    VB_TRY(make_params(b, estimator, gravity, max_dist, &pp, &sp));
    sp.rel_fitness = rel_fitness;
    sp.rel_rmse = rel_rmse;
    sp.max_iter = max_iter;
    if (b->P == 0) return VB200_OK;
    double *totals = b->d_totals_ext ? b->d_totals_ext : b->d_totals;
    if (!totals) return VB200_ERR_INVALID;  // vb200_batch_pass has not run
    if (!b->d_npts_global) VB_CUDA(cudaMallocAsync((void **)&b->d_npts_global, sizeof(double) * (size_t)b->P, st));
    std::vector<double> np((size_t)b->P);
    for (int p = 0; p < b->P; p++) np[p] = npts_global ? (double)npts_global[p] : (double)b->probs[p].npts;
    VB_CUDA(cudaMemcpyAsync(b->d_npts_global, np.data(), sizeof(double) * (size_t)b->P, cudaMemcpyHostToDevice, st));
    k_solve_totals<<<div_up(b->P, 32), 32, 0, st>>>(b->P, b->d_states, totals, b->d_npts_global, sp, pass_index);
    b->launches++;
    VB_CUDA(cudaGetLastError());
    VB_CUDA(cudaStreamSynchronize(st));  // `np` is pageable host memory
    return VB200_OK;
}

// n unconditional iterations from the current transforms (no convergence test, `done` never set)
static int batch_iterate(Batch *b, int estimator, const double *gravity, double max_dist, int n_iter) {
    Scene *sc = b->scene;
    cudaStream_t st = b->stream;
    if (estimator < VB200_EST_P2P || estimator > VB200_EST_P2PLANE_GRAVITY || n_iter < 0) return VB200_ERR_INVALID;
    if (!(max_dist > 0.0)) return VB200_ERR_DISTANCE;
    if (max_dist > sc->grid.p.cell * (1.0 + 1e-12)) return VB200_ERR_INVALID;
    const bool plane = estimator != VB200_EST_P2P;
    if (plane && (!b->has_normals || !sc->has_normals)) return VB200_ERR_NORMALS;
    if (b->P == 0 || b->nblk == 0) return VB200_OK;
    const PassParams pp = make_pass_params(sc->grid.p, max_dist);
    SolveParams sp;
    sp.ndone = nullptr;
    sp.rel_fitness = sp.rel_rmse = -1.0;  // |delta| < -1 never holds: no convergence exit
    sp.max_iter = 0x7fffffff;
    sp.estimator = estimator;
    sp.g[0] = 0.0; sp.g[1] = 1.0; sp.g[2] = 0.0;
    if (estimator == VB200_EST_P2PLANE_GRAVITY && gravity) {
        double l = sqrt(gravity[0] * gravity[0] + gravity[1] * gravity[1] + gravity[2] * gravity[2]);
        if (!(l > 0.0)) return VB200_ERR_INVALID;
        for (int a = 0; a < 3; a++) sp.g[a] = gravity[a] / l;
    }
    for (int i = 0; i < 3; i++)
        if (!b->ev[i]) VB_CUDA(cudaEventCreate(&b->ev[i]));
    const bool timed = n_iter == 1;  // per-kernel events only make sense around a single iteration
    for (int it = 0; it < n_iter; it++) {
        if (timed) VB_CUDA(cudaEventRecord(b->ev[0], st));
        if (b->all_nonempty) {
            launch_pass(b, plane, pp, &sp, b->iter_base++);
            if (timed) VB_CUDA(cudaEventRecord(b->ev[1], st));
        } else {
            launch_pass(b, plane, pp);
            if (timed) VB_CUDA(cudaEventRecord(b->ev[1], st));
            k_solve<<<b->P, 256, 0, st>>>(b->d_probs, b->d_states, b->d_partials, b->d_hard_cnt, sp, b->iter_base++);
            b->launches += 1;
        }
        if (timed) VB_CUDA(cudaEventRecord(b->ev[2], st));
    }
    b->ev_valid = timed;
    VB_CUDA(cudaGetLastError());
    return VB200_OK;
}

constexpr int64_t kPipelineMinPoints = 200000;  // below this a second stream costs more than the copy it hides

// the scene's second stream, created on first use
static bool second_stream(Scene *sc) {
    if (sc->stream2) return true;
    if (cudaStreamCreateWithFlags(&sc->stream2, cudaStreamNonBlocking) != cudaSuccess) {
        (void)cudaGetLastError();
        sc->stream2 = nullptr;
        return false;
    }
    return true;
}

}  // namespace vb

// ======================================================================================================
using vb::Batch;
using vb::Scene;

extern "C" int vb200_batch_iterate(vb200_batch_t *batch, int estimator, const double *gravity_axis,
                                   double max_dist, int n_iter) {
    if (!batch) return VB200_ERR_INVALID;
    Batch *b = reinterpret_cast<Batch *>(batch);
    VB_CUDA(cudaSetDevice(b->scene->device));
    return vb::batch_iterate(b, estimator, gravity_axis, max_dist, n_iter);
}

extern "C" int vb200_batch_pass(vb200_batch_t *batch, int estimator, double max_dist) {
    if (!batch) return VB200_ERR_INVALID;
    Batch *b = reinterpret_cast<Batch *>(batch);
    VB_CUDA(cudaSetDevice(b->scene->device));
    return vb::batch_pass(b, estimator, max_dist);
}

extern "C" int vb200_batch_set_totals_buffer(vb200_batch_t *batch, void *d_totals) {
    if (!batch) return VB200_ERR_INVALID;
    reinterpret_cast<Batch *>(batch)->d_totals_ext = (double *)d_totals;
    return VB200_OK;
}

extern "C" void *vb200_batch_totals(vb200_batch_t *batch) {
    if (!batch) return nullptr;
    Batch *b = reinterpret_cast<Batch *>(batch);
    return b->d_totals_ext ? b->d_totals_ext : b->d_totals;
}

extern "C" int vb200_batch_solve(vb200_batch_t *batch, int estimator, const double *gravity_axis, double max_dist,
                                 double rel_fitness, double rel_rmse, int max_iter, int pass_index,
                                 const int64_t *npts_global) {
    if (!batch || pass_index < 0 || max_iter < 0) return VB200_ERR_INVALID;
    Batch *b = reinterpret_cast<Batch *>(batch);
    VB_CUDA(cudaSetDevice(b->scene->device));
    return vb::batch_solve(b, estimator, gravity_axis, max_dist, rel_fitness, rel_rmse, max_iter, pass_index,
                           npts_global);
}

extern "C" int vb200_batch_last_kernel_ms(vb200_batch_t *batch, float *pass_ms, float *solve_ms) {
    if (!batch) return VB200_ERR_INVALID;
    Batch *b = reinterpret_cast<Batch *>(batch);
    if (!b->ev_valid) return VB200_ERR_INVALID;
    VB_CUDA(cudaSetDevice(b->scene->device));
    VB_CUDA(cudaEventSynchronize(b->ev[2]));
    float a = 0.f, c = 0.f;
    VB_CUDA(cudaEventElapsedTime(&a, b->ev[0], b->ev[1]));
    VB_CUDA(cudaEventElapsedTime(&c, b->ev[1], b->ev[2]));
    if (pass_ms) *pass_ms = a;
    if (solve_ms) *solve_ms = c;
    return VB200_OK;
}

static int batch_create_on(vb200_scene_t *scene, cudaStream_t stream, const double *src_xyz, const double *src_nrm,
                           const int64_t *src_offsets, int32_t n_clouds, vb200_batch_t **out);

extern "C" int vb200_batch_create(vb200_scene_t *scene, const double *src_xyz, const double *src_nrm,
                                  const int64_t *src_offsets, int32_t n_clouds, vb200_batch_t **out) {
    return batch_create_on(scene, nullptr, src_xyz, src_nrm, src_offsets, n_clouds, out);
}

static int batch_create_on(vb200_scene_t *scene, cudaStream_t stream, const double *src_xyz, const double *src_nrm,
                           const int64_t *src_offsets, int32_t n_clouds, vb200_batch_t **out) {
    if (!out) return VB200_ERR_INVALID;
    *out = nullptr;
    if (!scene || !src_offsets || n_clouds < 0 || (!src_xyz && src_offsets[n_clouds] > src_offsets[0]))
        return VB200_ERR_INVALID;
    Scene *sc = reinterpret_cast<Scene *>(scene);
    VB_CUDA(cudaSetDevice(sc->device));
    Batch *b = new Batch();
    b->scene = sc;
    b->stream = stream ? stream : sc->stream;
    b->has_normals = src_nrm != nullptr;  // only presence matters (Registration.cpp:152-157)
    int rc = vb::batch_upload(b, src_xyz, src_offsets, n_clouds);
    if (rc != VB200_OK) {
        vb::batch_free(b);
        return rc;
    }
    *out = reinterpret_cast<vb200_batch_t *>(b);
    return VB200_OK;
}

extern "C" int vb200_batch_destroy(vb200_batch_t *batch) {
    if (!batch) return VB200_OK;
    Batch *b = reinterpret_cast<Batch *>(batch);
    cudaSetDevice(b->scene->device);
    cudaStreamSynchronize(b->stream);
    vb::batch_free(b);
    return VB200_OK;
}

extern "C" int vb200_batch_set_problems(vb200_batch_t *batch, const int32_t *cloud_ids, const double *init_T,
                                        int32_t P) {
    if (!batch || P < 0 || (P > 0 && !init_T)) return VB200_ERR_INVALID;
    Batch *b = reinterpret_cast<Batch *>(batch);
    VB_CUDA(cudaSetDevice(b->scene->device));
    return vb::batch_set_problems(b, cloud_ids, init_T, P);
}

extern "C" int vb200_batch_run(vb200_batch_t *batch, int estimator, const double *gravity_axis, double max_dist,
                               double rel_fitness, double rel_rmse, int max_iter) {
    if (!batch) return VB200_ERR_INVALID;
    Batch *b = reinterpret_cast<Batch *>(batch);
    VB_CUDA(cudaSetDevice(b->scene->device));
    return vb::batch_run(b, estimator, gravity_axis, max_dist, rel_fitness, rel_rmse, max_iter);
}

extern "C" int vb200_batch_results(vb200_batch_t *batch, double *out_T, double *out_fitness, double *out_rmse,
                                   int32_t *out_ncorr, int32_t *out_iters) {
    if (!batch) return VB200_ERR_INVALID;
    Batch *b = reinterpret_cast<Batch *>(batch);
    VB_CUDA(cudaSetDevice(b->scene->device));
    std::vector<vb::ProbState> states((size_t)b->P);
    if (b->P)
        VB_CUDA(cudaMemcpyAsync(states.data(), b->d_states, sizeof(vb::ProbState) * (size_t)b->P,
                                cudaMemcpyDeviceToHost, b->stream));
    VB_CUDA(cudaStreamSynchronize(b->stream));
    for (int p = 0; p < b->P; p++) {
        const vb::ProbState &s = states[p];
        if (out_T) memcpy(out_T + 16 * (size_t)p, s.T, sizeof(double) * 16);
        if (out_fitness) out_fitness[p] = s.fitness;
        if (out_rmse) out_rmse[p] = s.rmse;
        if (out_ncorr) out_ncorr[p] = s.ncorr;
        if (out_iters) out_iters[p] = s.iters;
    }
    return VB200_OK;
}

extern "C" int vb200_batch_corr(vb200_batch_t *batch, int32_t p, int32_t *out_corr, int32_t *out_k) {
    if (!batch || !out_corr || !out_k) return VB200_ERR_INVALID;
    Batch *b = reinterpret_cast<Batch *>(batch);
    if (p < 0 || p >= b->P) return VB200_ERR_INVALID;
    VB_CUDA(cudaSetDevice(b->scene->device));
    const vb::ProbDesc &pd = b->probs[p];
    std::vector<int> cj((size_t)pd.npts), so((size_t)pd.npts);
    vb::DevBuf<int> d_j(b->stream);
    if (pd.npts) {
        VB_CUDA(d_j.alloc((size_t)pd.npts));
        vb::k_corr_orig<<<vb::div_up(pd.npts, 256), 256, 0, b->stream>>>(b->d_corr + pd.corr_begin, pd.npts,
                                                                               b->scene->grid.orig, d_j.p);
        VB_CUDA(cudaGetLastError());
        VB_CUDA(cudaMemcpyAsync(cj.data(), d_j.p, sizeof(int) * (size_t)pd.npts,
                                cudaMemcpyDeviceToHost, b->stream));
        VB_CUDA(cudaMemcpyAsync(so.data(), b->d_src_orig + pd.src_begin, sizeof(int) * (size_t)pd.npts,
                                cudaMemcpyDeviceToHost, b->stream));
    }
    VB_CUDA(cudaStreamSynchronize(b->stream));
    // sorted position -> original source index, then emit in ascending source index
    std::vector<int> by_src((size_t)pd.npts, -1);
    for (int s = 0; s < pd.npts; s++) by_src[(size_t)so[s]] = cj[s];
    int k = 0;
    for (int i = 0; i < pd.npts; i++)
        if (by_src[i] >= 0) {
            out_corr[2 * k] = i;
            out_corr[2 * k + 1] = by_src[i];
            k++;
        }
    *out_k = k;
    return VB200_OK;
}

extern "C" int64_t vb200_batch_launches(const vb200_batch_t *batch) {
    return batch ? reinterpret_cast<const Batch *>(batch)->launches : 0;
}

extern "C" int vb200_icp_run(vb200_scene_t *scene, const double *src_xyz, const double *src_nrm,
                             const int64_t *src_offsets, int32_t B, const double *init_T, int estimator,
                             const double *gravity_axis, double max_dist, double rel_fitness, double rel_rmse,
                             int max_iter, double *out_T, double *out_fitness, double *out_rmse,
                             int32_t *out_ncorr, int32_t *out_iters, int32_t *out_corr) {
    if (!scene || !src_offsets || B < 0 || (B > 0 && !init_T)) return VB200_ERR_INVALID;
    // the reference's early-outs return RegistrationResult(init): transformation = init, fitness = rmse = 0
    auto passthrough = [&]() {
        for (int p = 0; p < B; p++) {
            if (out_T) memcpy(out_T + 16 * (size_t)p, init_T + 16 * (size_t)p, sizeof(double) * 16);
            if (out_fitness) out_fitness[p] = 0.0;
            if (out_rmse) out_rmse[p] = 0.0;
            if (out_ncorr) out_ncorr[p] = 0;
            if (out_iters) out_iters[p] = 0;
        }
    };
    Scene *sc = reinterpret_cast<Scene *>(scene);
    if (!(max_dist > 0.0)) { passthrough(); return VB200_ERR_DISTANCE; }
    if (estimator != VB200_EST_P2P && (!src_nrm || !sc->has_normals)) { passthrough(); return VB200_ERR_NORMALS; }
    // Two halves on two streams when there is enough to move: the second half's host-to-device copy and
    // spatial sort run while the first half's opening iterations compute (the objects are independent, so
    // nothing else changes).  38 MB of sources are ~0.8 ms of PCIe time, half of which this hides.
    VB_CUDA(cudaSetDevice(sc->device));
    const int64_t total = src_offsets[B] - src_offsets[0];
    int split = B;  // clouds [0, split) on the scene's stream, [split, B) on its second one
    if (B >= 2 && total >= vb::kPipelineMinPoints && max_iter >= 4 && vb::second_stream(sc)) {
        split = 1;
        while (split < B - 1 && 2 * (src_offsets[split] - src_offsets[0]) < total) split++;
    }
    vb200_batch_t *half[2] = {nullptr, nullptr};
    const int first[3] = {0, split, B};
    const int nhalf = split < B ? 2 : 1;
    const int warm = 3;  // iterations 0..3 of the first half are enqueued before the second half is uploaded
    int rc = VB200_OK;
    for (int h = 0; h < nhalf && rc == VB200_OK; h++) {
        const int p0 = first[h], np = first[h + 1] - first[h];
        rc = batch_create_on(scene, h == 0 ? nullptr : sc->stream2, src_xyz, src_nrm, src_offsets + p0, np, &half[h]);
        if (rc == VB200_OK) rc = vb200_batch_set_problems(half[h], nullptr, init_T + 16 * (size_t)p0, np);
        if (rc == VB200_OK)
            rc = vb::batch_run(reinterpret_cast<Batch *>(half[h]), estimator, gravity_axis, max_dist, rel_fitness,
                               rel_rmse, max_iter, 0, nhalf == 2 ? warm : 0x7fffffff);
    }
    for (int h = 0; h < nhalf && rc == VB200_OK && nhalf == 2; h++)
        rc = vb::batch_run(reinterpret_cast<Batch *>(half[h]), estimator, gravity_axis, max_dist, rel_fitness, rel_rmse,
                           max_iter, warm + 1);
    for (int h = 0; h < nhalf && rc == VB200_OK; h++) {
        const int p0 = first[h], np = first[h + 1] - first[h];
        rc = vb200_batch_results(half[h], out_T ? out_T + 16 * (size_t)p0 : nullptr, out_fitness ? out_fitness + p0 : nullptr,
                                 out_rmse ? out_rmse + p0 : nullptr, out_ncorr ? out_ncorr + p0 : nullptr,
                                 out_iters ? out_iters + p0 : nullptr);
        if (rc == VB200_OK && out_corr) {
            for (int p = 0; p < np && rc == VB200_OK; p++) {
                int32_t k = 0;
                rc = vb200_batch_corr(half[h], p, out_corr + 2 * (src_offsets[p0 + p] - src_offsets[0]), &k);
            }
        }
    }
    for (int h = 0; h < 2; h++)
        if (half[h]) vb200_batch_destroy(half[h]);
    return rc;
}

extern "C" int vb200_register_model_to_scene(vb200_scene_t *scan, const double *model_xyz, const double *model_nrm,
                                             int64_t m, int level, double threshold, int point_to_plane,
                                             double out_T[16], int32_t *out_ncorr, int32_t *out_best_level) {
    if (!scan || !model_xyz || m < 0 || level <= 0 || !out_T) return VB200_ERR_INVALID;
    const int64_t off[2] = {0, m};
    vb200_batch_t *batch = nullptr;
    int rc = vb200_batch_create(scan, model_xyz, model_nrm, off, 1, &batch);
    if (rc != VB200_OK) return rc;
    std::vector<double> inits(16 * (size_t)level, 0.0);
    std::vector<int32_t> ids((size_t)level, 0);
    const double interval = 2 * M_PI / level;  // src/annotation.cpp:35
    for (int i = 0; i < level; i++) {
        double a = interval * i, c = cos(a), s = sin(a);
        double *T = inits.data() + 16 * (size_t)i;  // AngleAxis(a, UnitY) (src/annotation.cpp:41-43)
        T[0] = c; T[2] = s; T[5] = 1.0; T[8] = -s; T[10] = c; T[15] = 1.0;
    }
    rc = vb200_batch_set_problems(batch, ids.data(), inits.data(), level);
    // ICPConvergenceCriteria() defaults: 1e-6, 1e-6, 30 (Registration.h:49-50)
    if (rc == VB200_OK)
        rc = vb200_batch_run(batch, point_to_plane ? VB200_EST_P2PLANE : VB200_EST_P2P, nullptr, threshold, 1e-6, 1e-6, 30);
    std::vector<double> Ts(16 * (size_t)level);
    std::vector<int32_t> nc((size_t)level);
    if (rc == VB200_OK) rc = vb200_batch_results(batch, Ts.data(), nullptr, nullptr, nc.data(), nullptr);
    vb200_batch_destroy(batch);
    if (rc == VB200_ERR_DISTANCE || rc == VB200_ERR_NORMALS) {
        // every RegistrationICP call returned RegistrationResult(init) with no correspondences, so
        // best_result stays default-constructed: Identity (src/annotation.cpp:36,59-63)
        for (int i = 0; i < 16; i++) out_T[i] = (i % 5 == 0) ? 1.0 : 0.0;
        if (out_ncorr) *out_ncorr = 0;
        if (out_best_level) *out_best_level = -1;
        return rc;
    }
    if (rc != VB200_OK) return rc;
    int best = -1, best_k = 0;
    for (int i = 0; i < level; i++)
        if (nc[i] > best_k) { best_k = nc[i]; best = i; }  // strict >, first wins (src/annotation.cpp:59)
    for (int i = 0; i < 16; i++) out_T[i] = best >= 0 ? Ts[16 * (size_t)best + i] : ((i % 5 == 0) ? 1.0 : 0.0);
    if (out_ncorr) *out_ncorr = best_k;
    if (out_best_level) *out_best_level = best;
    return VB200_OK;
}

extern "C" int vb200_estimate(const double *src_xyz, int64_t m, const double *tgt_xyz, const double *tgt_nrm,
                              int64_t n, const int32_t *corr, int64_t K, int estimator, const double *gravity_axis,
                              int device, double out_T[16]) {
    if (!out_T || K < 0 || (K > 0 && (!src_xyz || !tgt_xyz || !corr))) return VB200_ERR_INVALID;
    if (estimator < VB200_EST_P2P || estimator > VB200_EST_P2PLANE_GRAVITY) return VB200_ERR_INVALID;
    for (int i = 0; i < 16; i++) out_T[i] = (i % 5 == 0) ? 1.0 : 0.0;
    const bool plane = estimator != VB200_EST_P2P;
    // corres.empty() || !target.HasNormals() -> Identity (TransformationEstimation.cpp:51,79-80)
    if (K == 0 || (plane && !tgt_nrm)) return VB200_OK;
    VB_TRY(vb::select_device(device));
    // marshal the K referenced rows (the reference gathers them too, TransformationEstimation.cpp:52-57)
    std::vector<double> h((size_t)K * 9);
    double *vs = h.data(), *vt = vs + 3 * K, *nt = vt + 3 * K;
    for (int64_t i = 0; i < K; i++) {
        int64_t a = corr[2 * i], bq = corr[2 * i + 1];
        if (a < 0 || a >= m || bq < 0 || bq >= n) return VB200_ERR_INVALID;
        for (int c = 0; c < 3; c++) {
            vs[3 * i + c] = src_xyz[3 * a + c];
            vt[3 * i + c] = tgt_xyz[3 * bq + c];
            nt[3 * i + c] = plane ? tgt_nrm[3 * bq + c] : 0.0;
        }
    }
    vb::SolveParams sp;
    sp.ndone = nullptr;
    sp.rel_fitness = sp.rel_rmse = 0.0;
    sp.max_iter = 0;
    sp.estimator = estimator;
    sp.g[0] = 0.0; sp.g[1] = 1.0; sp.g[2] = 0.0;
    if (estimator == VB200_EST_P2PLANE_GRAVITY && gravity_axis) {
        double l = sqrt(gravity_axis[0] * gravity_axis[0] + gravity_axis[1] * gravity_axis[1] + gravity_axis[2] * gravity_axis[2]);
        if (!(l > 0.0)) return VB200_ERR_INVALID;
        for (int a = 0; a < 3; a++) sp.g[a] = gravity_axis[a] / l;
    }
    const int nblk = vb::div_up(K, vb::kChunk);
    vb::DevBuf<double> d_in, d_part, d_T;
    VB_CUDA(d_in.alloc((size_t)K * 9));
    VB_CUDA(d_part.alloc((size_t)nblk * vb::kAcc));
    VB_CUDA(d_T.alloc(16));
    VB_CUDA(cudaMemcpy(d_in.p, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice));
    if (plane)
        vb::k_estimate<1><<<nblk, vb::kPassTpb>>>(d_in.p, d_in.p + 3 * K, d_in.p + 6 * K, K, d_part.p);
    else
        vb::k_estimate<0><<<nblk, vb::kPassTpb>>>(d_in.p, d_in.p + 3 * K, d_in.p + 6 * K, K, d_part.p);
    vb::k_estimate_solve<<<1, 256>>>(d_part.p, nblk, d_in.p, sp, d_T.p);
    VB_CUDA(cudaGetLastError());
    VB_CUDA(cudaMemcpy(out_T, d_T.p, sizeof(double) * 16, cudaMemcpyDeviceToHost));
    return VB200_OK;
}

#ifdef VB_STATS
// dev builds only (scripts/build_variants.sh ... "-DVB_STATS"): search-path counters of grid.cuh
extern "C" int vb200_debug_stats(unsigned long long *out, int reset) {
    if (out) cudaMemcpyFromSymbol(out, vb::g_stats, sizeof(unsigned long long) * 16);
    if (reset) {
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(vb::g_stats, z, sizeof(z));
    }
    return 0;
}
#endif
