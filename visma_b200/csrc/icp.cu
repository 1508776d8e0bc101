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

#ifndef VB_PASS_TPB
#define VB_PASS_TPB 128
#endif
#ifndef VB_PASS_B_TPB
#define VB_PASS_B_TPB 1024  // part B's block: independent warps (no block-level barrier).  One block per SM: when nothing is
                            // listed the launch is 148 blocks to start and retire instead of 1 184 (a settled iteration of one
                            // rank's 4 objects: 33 -> 30 us), and the search-bound passes are 2-3 % faster than with 8 x 128
#endif
constexpr int kPassBTpb = VB_PASS_B_TPB;
constexpr int kPassBWarps = kPassBTpb / 32;
#ifndef VB_PASS_B_WARPS_PER_SM
#define VB_PASS_B_WARPS_PER_SM 32  // 64 registers: the search is latency-bound, 8 more warps per SM hide more than the extra spills cost
#endif
constexpr int kPassBBlocksPerSM = VB_PASS_B_WARPS_PER_SM * 32 / kPassBTpb;
constexpr int kPassTpb = VB_PASS_TPB;
constexpr int kPassWarps = kPassTpb / 32;  // every warp writes its own partial: no block-level barrier
#ifndef VB_PASS_PTS
#define VB_PASS_PTS 4
#endif
constexpr int kPtsPerThread = VB_PASS_PTS;  // the most a part-A thread takes (Batch::pts picks 4, 2 or 1 per batch of problems)
constexpr int kChunk = kPassTpb * kPtsPerThread;  // the most source points a block covers
constexpr int kAcc = 32;                          // accumulator slots (padded)
constexpr int kBuckets = 32768;                   // spatial buckets per source cloud (15-bit key)

// accumulator slots, MODE 1 (point-to-plane): 0..20 JTJ upper triangle row-major, 21..26 JTr
// MODE 0 (point-to-point): 0..2 sum s', 3..5 sum d', 6..14 sum d' s'^T (row-major), 15 sum |s'|^2
// both: 28 = sum |s - d|^2 (estimator plug-in only), 30 = sum d2 (exact NN distances), 31 = count
constexpr int kSlotRes2 = 28, kSlotD2 = 30, kSlotCount = 31;

struct ProbState {
    double T[16];
    double fitness, rmse, prev_fitness, prev_rmse;
    int ncorr, iters, done;
    // How far ANY point of the problem can have moved, as an upper bound accumulated over the estimator updates
    // since set_problems (k_solve adds each update's bound, rounded up): the difference between two readings
    // bounds the move of every point between them, which is all the cached-neighbour tests need to know.
    float cum;
    float last_move;  // the latest update's share (huge before the first update)
    // the cloud's bounding sphere under the current transform lies inside the grid (k_solve's verdict, rigorous, 0 until
    // the first update): part A then skips the per-point range test
    int all_inside;
};

struct BlockTask {
    int prob;       // problem index
    int src_begin;  // first source point (global sorted position)
    int count;      // points in this block's chunk
    int corr_begin; // where this chunk's records start
};

struct ProbDesc {
    int cloud;
    int npts;
    int blk_begin, blk_count;
    int src_begin;
    int corr_begin;
    double ctr[3];  // centre of the cloud's bounding box (its own frame) ...
    double rad;     // ... and the largest distance of a point from it
};

// What a source point remembers between passes, per (problem, point).  `hot` is all a settled pass reads:
//   c0    the point's match as a SORTED scene position (-1 none): the correspondence list, and the next search's bound;
//   lim1  sqrt(sec1) + cum at the time of the proof, rounded down, where sec1 bounds (from below) the true squared
//         distance from the point's position THEN to every scene point other than c0.  While
//         |q_now - c0| + cum_now < lim1 the triangle inequality proves c0 is still the unique nearest neighbour.
//         NaN (memset 0xff) or 0 = knows nothing.
// `cold` is read only when that test fails: three more candidates and limK, the same bound for every scene point
// outside {c0, c1, c2, c3}: while min_j |q_now - c_j| + cum_now < limK the nearest neighbour is one of the four and
// is settled among them in double.
struct __align__(8) HotRec {
    int c0;
    float lim1;
};
struct __align__(16) ColdRec {
    int c1, c2, c3;
    float limK;
};

struct PassParams {
    double r2;        // (double)(float)(max_dist^2)  (KDTreeFlann.cpp:185)
    float r2_ub;      // f32 upper bound of r2 including the screening band
    float pos_err;    // absolute guard of the cached-neighbour tests (f32 centring of the query, f64 rounding of T p)
    float slack_lo, slack_hi;  // how far beyond the answer's reach a search of a settling problem looks (metric)
    float set_move;   // a problem whose latest update moved it by less than this is "settling": its searches keep sets
    int use_cache;    // 0: every point is searched in every pass (the ablation bench.py reports)
    int *work;        // part B's work items of this pass — (block << 4 | batch) — in order of arrival
    int *work_ctr;    // [parity][2]: {items listed, items handed out}; passes alternate between the two pairs
    int parity;
    double *brows;    // per (block, batch): the batch's Gram matrix, until the block's last batch folds them
    int *blk_done;    // per block: batches finished in this pass (reset by the one that folds)
    int64_t src_n;    // points in the batch's source array (x[src_n], y[src_n], z[src_n])
    int pts;          // points per part-A thread: a block covers kPassTpb * pts consecutive points (see Batch::pts)
};

#ifndef VB_SLACK_LO_PCT
#define VB_SLACK_LO_PCT 1
#endif
#ifndef VB_SLACK_HI_PCT
#define VB_SLACK_HI_PCT 20
#endif
#ifndef VB_SET_MOVE_PCT
#define VB_SET_MOVE_PCT 20
#endif
#ifndef VB_SLACK_MOVE_PCT
#define VB_SLACK_MOVE_PCT 100  // what is still to come moves a point by less than the last update did (updates shrink ~3-4x per
                               // iteration), and its distance to the match can grow by as much again
#endif
inline PassParams make_pass_params(const GridParams &g, double max_dist, bool use_cache) {
    PassParams pp;
    memset(&pp, 0, sizeof(pp));
    pp.r2 = (double)(float)(max_dist * max_dist);
    pp.r2_ub = r2_upper_bound(g, pp.r2);
    pp.slack_lo = g.fine * (VB_SLACK_LO_PCT * 0.01f);
    pp.slack_hi = g.fine * (VB_SLACK_HI_PCT * 0.01f);
    pp.set_move = g.fine * (VB_SET_MOVE_PCT * 0.01f);
    pp.pos_err = g.band_a;  // band_a = 2 sqrt(3) e with e the per-axis f32 rounding of a centred coordinate
    pp.use_cache = use_cache ? 1 : 0;
    return pp;
}

struct SolveParams {
    double rel_fitness, rel_rmse;
    double g[3];
    int max_iter;
    int estimator;
    int *ndone;  // nullable: counts the problems that have finished (lets the host stop enqueueing iterations)
    double box_lo[3], box_hi[3];  // the scene grid's box, shrunk by a rounding margin (for ProbState::all_inside)
};

// ---- programmatic dependent launch: the three kernels of an iteration are chained with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the next kernel's blocks are resident and past their
// prologue when the previous one drains.  Nothing produced by the previous kernel is touched before pdl_wait().
// One-sided fences at GPU scope.  __threadfence() is fence.sc: MEMBAR.SC + CCTL.IVALL, i.e. it also throws away every
// line of the SM's L1 — the release side of a hand-off needs no invalidation at all, the acquire side only that.
__device__ __forceinline__ void fence_release_gpu() { asm volatile("fence.release.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_acquire_gpu() { asm volatile("fence.acquire.gpu;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }

// ---- warp-level vector reduction: 32 slots over 32 lanes in 31 exchange steps (recursive halving).
// After the call lane L holds the warp total of slot L.  Fixed order => deterministic.  (estimator plug-in)
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
__device__ __forceinline__ void contributions(bool matched, const double *vs, const double *vt,
                                              const double *nt, const double *cref, double (&v)[kAcc]) {
#pragma unroll
    for (int j = 0; j < kAcc; j++) v[j] = 0.0;
    if (!matched) return;
    const double ex = vs[0] - vt[0], ey = vs[1] - vt[1], ez = vs[2] - vt[2];
    if (MODE == 1) {
        // r = (vs - vt).nt ; J = [vs x nt ; nt]  (TransformationEstimation.cpp:87-89)
        double r = ex * nt[0] + ey * nt[1] + ez * nt[2];
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
    v[kSlotRes2] = (ex * ex + ey * ey) + ez * ez;  // (s - t).squaredNorm()  (src/constrained_ICP.cpp:19)
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
static_assert(kPassTpb == kPassWarps * 32 && kChunk / 32 <= 16, "work items are (block << 4 | batch)");

// D(8x8) += A(8x4) * B(4x8) in FP64 on the tensor cores (DMMA).  Lane l holds A[l>>2][l&3], B[l&3][l>>2] and
// D[l>>2][2(l&3) + {0,1}].
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// One ICP correspondence pass for every active problem: transform + radius-bounded 1-NN + estimator
// products + reduction.  grid = one block per kChunk source points of one problem; every warp owns
// kPtsPerThread batches of 32 consecutive (spatially sorted) points.
//
// 1. Cached-neighbour tests (Greenspan & Godin 2001, made exact; see HotRec): a point whose last search left a
//    bound on everything but its match (or on everything but a set of four candidates) needs no search while the
//    problem has moved by less than that bound allows.  Once an alignment settles this is nearly every point, and
//    the pass becomes a streaming gather/reduce.
// 2. The points that fail are compacted per warp (ballot ranks, deterministic) and searched 32 at a time with all
//    lanes busy (k_pass_b_wl, nn_search_hybrid); each search refreshes the point's records.
// The decision is always taken in double from the exact coordinates: d2 = |q - b|^2 with FLANN's operation
// order, accepted iff d2 < (double)(float)(r*r).
//
// Estimator sums: every matched lane stages one row x of 8 doubles — point-to-plane [J (6), r, 0] with
// r = (vs - vt).nt, J = [vs x nt ; nt] (TransformationEstimation.cpp:87-89); point-to-point
// [s' (3), d' (3), 1, 0] with (s', d') = (vs - c, vt - c) — and the warp accumulates the Gram matrix
// sum_i x_i x_i^T with 8 DMMA instructions per 32 rows: JTJ = D[0..5][0..5], JTr = D[0..5][6]
// (resp. the umeyama moments sum d' s'^T = D[3..5][0..2], sum s' = D[0..2][6], sum d' = D[3..5][6]).  The
// accumulator fragment lives in two registers per lane for the whole warp.
template <int MODE>
struct PassCtx {
    const GridDev &G;
    const double *T;  // the problem's current transform (global memory, warp-uniform loads)
    double c0 = 0.0, c1 = 0.0;  // this lane's two entries of the warp's Gram matrix
    double sum_d2 = 0.0;        // exact NN distances (Registration.cpp:68), this lane's share
    int count = 0;

    __device__ __forceinline__ PassCtx(const GridDev &g, const double *t) : G(g), T(t) {}

    // source point i of the batch's sorted clouds (x[n], y[n], z[n])
    __device__ __forceinline__ void transform(const double *__restrict__ src, int64_t n, int64_t i, double (&vs)[3]) const {
        const double px = src[i], py = src[n + i], pz = src[2 * n + i];
        vs[0] = T[0] * px + T[1] * py + T[2] * pz + T[3];
        vs[1] = T[4] * px + T[5] * py + T[6] * pz + T[7];
        vs[2] = T[8] * px + T[9] * py + T[10] * pz + T[11];
    }

    // the exact decision for a match candidate at vt (normal nt) and, if accepted, the lane's estimator row
    __device__ __forceinline__ bool row_from(const double (&vs)[3], const double (&vt)[3], const double (&nt)[3], double r2,
                                             double (&x)[8]) {
        const double d2 = l2_exact(vs[0], vs[1], vs[2], vt[0], vt[1], vt[2]);
        if (!(d2 < r2)) return false;
        fill_row(vs, vt, nt, d2, x);
        return true;
    }
    // the accepted match's estimator row and its share of the pass's sums
    __device__ __forceinline__ void fill_row(const double (&vs)[3], const double (&vt)[3], const double (&nt)[3], double d2,
                                             double (&x)[8]) {
        if (MODE == 1) {
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
    }
    // ... for match candidate bs (>= 0), gathered here: one 256-bit load per record
    __device__ __forceinline__ bool row_of(int bs, const double (&vs)[3], double r2, double (&x)[8]) {
        const Rec3 t = ld_rec(G.xyz + kPtStride * (int64_t)bs);
        const double vt[3] = {t.x, t.y, t.z};
        double nt[3] = {0.0, 0.0, 0.0};
        if (MODE == 1) {
            const Rec3 nn = ld_rec(G.nrm + kPtStride * (int64_t)bs);
            nt[0] = nn.x; nt[1] = nn.y; nt[2] = nn.z;
        }
        return row_from(vs, vt, nt, r2, x);
    }

    // Part A's form: the lane's row goes straight into its slot of the warp's staging area (no 8-double row held in
    // registers across the point's tests: at 40 registers it lived in local memory), zeros for a lane without a match.
    __device__ __forceinline__ void stage_row(double *rows, const double (&vs)[3], const double (&vt)[3], const double (&nt)[3],
                                              double d2) {
        double x[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        fill_row(vs, vt, nt, d2, x);
        double2 *row = reinterpret_cast<double2 *>(rows + (threadIdx.x & 31) * kRowStride);
        row[0] = make_double2(x[0], x[1]);
        row[1] = make_double2(x[2], x[3]);
        row[2] = make_double2(x[4], x[5]);
        row[3] = make_double2(x[6], x[7]);
    }
    __device__ __forceinline__ void stage_zero_row(double *rows) {
        double2 *row = reinterpret_cast<double2 *>(rows + (threadIdx.x & 31) * kRowStride);
        row[0] = row[1] = row[2] = row[3] = make_double2(0.0, 0.0);
    }
    // all 32 lanes, rows staged (the previous batch's reads ended with a __syncwarp)
    __device__ __forceinline__ void accumulate_staged(const double *rows) {
        const int lane = threadIdx.x & 31;
        __syncwarp();
#pragma unroll
        for (int ch = 0; ch < 8; ch++) {
            const double a = rows[(4 * ch + (lane & 3)) * kRowStride + (lane >> 2)];
            dmma_m8n8k4(c0, c1, a, a);
        }
        __syncwarp();  // rows consumed before the next batch's are written
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
constexpr int kBatchesPerBlock = kChunk / 32;  // a block's listed points form at most this many 32-point batches

// sqrt(x) * (1 - 1e-6) + cum, every step rounded DOWN: the stored side of a cached-neighbour test
__device__ __forceinline__ float lim_of(float sec, float cum) {
    if (!(sec > 0.0f)) return 0.0f;
    return __fadd_rd(__fmul_rd(__fsqrt_rd(sec), 0.999999f), cum);
}
// sqrt(d) * (1 + 1e-6) + cum, every step rounded UP: the moving side of the test (d = f32 distance + its band)
__device__ __forceinline__ float reach_now(float d_ub, float cum_eps) {
    return __fadd_ru(__fmul_ru(__fsqrt_ru(d_ub), 1.000001f), cum_eps);
}

// the same test without the square root: sqrt(d)(1 + 1e-6) + cum_eps < lim  <=>  d (1 + 1e-6)^2 < (lim - cum_eps)^2 for a
// positive right-hand side; left side rounded up, right side down.  A NaN or zero limit compares false.
__device__ __forceinline__ bool within_limit(float d_ub, float cum_eps, float lim) {
    const float m = __fsub_rd(lim, cum_eps);
    return m > 0.0f && __fmul_ru(d_ub, 1.0000021f) < __fmul_rd(m, m);
}

// Part A's rare path: the nearest neighbour is known to be one of cand[0..3] (-1 = unused) but their f32
// distances cannot order them.  Exact distances; ties to the lowest ORIGINAL index (the rule of nn_exact_rescan).
static __device__ __noinline__ int settle_set_exact(const GridDev &G, double qx, double qy, double qz, int c0, int c1, int c2,
                                                    int c3) {
    const int cand[4] = {c0, c1, c2, c3};
    int w = -1, wo = 0x7fffffff;
    double dw = 0.0;
    for (int j = 0; j < 4; j++) {
        if (cand[j] < 0) continue;
        const double d = l2_exact(qx, qy, qz, G.xyz + kPtStride * (int64_t)cand[j]);
        const int o = __ldg(G.orig + cand[j]);
        if (w < 0 || d < dw || (d == dw && o < wo)) { w = j; dw = d; wo = o; }
    }
    return w;
}

// Part A's second chance for a point whose single-candidate test failed: its candidate set (ColdRec).  Is the
// nearest of the four nearer than anything outside the set can be?  Then the nearest neighbour is that member
// (settled in double when f32 cannot order them); the records are refreshed.  Returns the match or -1.
static __device__ __forceinline__ int settle_from_set(const GridDev &G, const PassParams &pp, double qx, double qy, double qz,
                                                   int c0, int slot, float cum, float cum_eps, HotRec *__restrict__ hot,
                                                   ColdRec *__restrict__ cold) {
    const ColdRec kc = cold[slot];
    VB_STAT(19, kc.c1 >= 0);
    if (!(kc.c1 >= 0 && cum_eps < kc.limK)) return -1;
    VB_STAT(14, 1);
    QueryCtx c;
    make_query(G.p, qx, qy, qz, c);
    const float4 t0 = __ldg(G.hi + c0), t1 = __ldg(G.hi + kc.c1);
    float4 t2 = make_float4(3.0e18f, 3.0e18f, 3.0e18f, 0.0f), t3 = t2;
    if (kc.c2 >= 0) t2 = __ldg(G.hi + kc.c2);
    if (kc.c3 >= 0) t3 = __ldg(G.hi + kc.c3);
    const float ax = c.qx - t0.x, ay = c.qy - t0.y, az = c.qz - t0.z;
    const float bx = c.qx - t1.x, by = c.qy - t1.y, bz = c.qz - t1.z;
    const float cx = c.qx - t2.x, cy = c.qy - t2.y, cz = c.qz - t2.z;
    const float ex = c.qx - t3.x, ey = c.qy - t3.y, ez = c.qz - t3.z;
    const float d0 = fmaf(az, az, fmaf(ay, ay, ax * ax));
    const float d1 = fmaf(bz, bz, fmaf(by, by, bx * bx));
    const float d2 = fmaf(cz, cz, fmaf(cy, cy, cx * cx));
    const float d3 = fmaf(ez, ez, fmaf(ey, ey, ex * ex));
    const float dmin = fminf(fminf(d0, d1), fminf(d2, d3));
    if (!(reach_now(dmin + band(G.p, dmin), cum_eps) < kc.limK)) return -1;
    VB_STAT(15, 1);
    // the f32-nearest member, and the nearest of the others
    int w = 0;
    float bw = d0;
    if (d1 < bw) { w = 1; bw = d1; }
    if (d2 < bw) { w = 2; bw = d2; }
    if (d3 < bw) { w = 3; bw = d3; }
    float rest = fminf(fminf(w == 0 ? 3.0e38f : d0, w == 1 ? 3.0e38f : d1), fminf(w == 2 ? 3.0e38f : d2, w == 3 ? 3.0e38f : d3));
    if (!(rest - bw > band(G.p, bw) + band(G.p, rest))) {
        w = settle_set_exact(G, qx, qy, qz, c0, kc.c1, kc.c2, kc.c3);
        rest = fminf(fminf(w == 0 ? 3.0e38f : d0, w == 1 ? 3.0e38f : d1), fminf(w == 2 ? 3.0e38f : d2, w == 3 ? 3.0e38f : d3));
    }
    const int match = w == 0 ? c0 : (w == 1 ? kc.c1 : (w == 2 ? kc.c2 : kc.c3));
    // A fresh single-candidate bound, valid as of NOW: the set's other members are at least sqrt(rest - band) away,
    // everything outside it at least limK - cum_now.
    const float others = fmaxf(rest - band(G.p, rest), 0.0f);
    const float outside = __fsub_rd(kc.limK, cum_eps);
    HotRec hn;
    hn.c0 = match;
    hn.lim1 = __fadd_rd(fminf(__fmul_rd(__fsqrt_rd(others), 0.999999f), outside), cum);
    hot[slot] = hn;
    if (w != 0) {
        ColdRec kn = kc;
        if (w == 1) kn.c1 = c0; else if (w == 2) kn.c2 = c0; else kn.c3 = c0;
        cold[slot] = kn;
    }
    return match;
}

// ---- pass, part A: every point.  Streaming: transform, cached-neighbour tests, and for the points they settle
// the exact decision and the estimator row.  The rest are listed per warp (ballot ranks: deterministic order)
// for part B.  No search code in here, so the kernel runs at high occupancy: it is a gather/reduce bound by
// memory latency and bandwidth.
template <int MODE>
__global__ void __launch_bounds__(kPassTpb, VB_PASS_A_MINBLOCKS) k_pass_a(
    GridDev G, const double *__restrict__ src_xyz, const BlockTask *__restrict__ tasks,
    const ProbState *__restrict__ states, double *__restrict__ partials, HotRec *__restrict__ hot,
    ColdRec *__restrict__ cold, unsigned char *__restrict__ hard_ids, int *__restrict__ hard_cnt, PassParams pp) {
    const BlockTask task = tasks[blockIdx.x];  // static since set_problems: safe before the dependency resolves
    pdl_launch_dependents();
    pdl_wait();  // the previous iteration's k_solve (T, cum) and part B (records) are complete and visible
    const ProbState *st = states + task.prob;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // the counters the NEXT pass will use: idle since the previous pass's part B finished
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        pp.work_ctr[2 * (pp.parity ^ 1)] = 0;
        pp.work_ctr[2 * (pp.parity ^ 1) + 1] = 0;
    }
    if (st->done) return;
    __shared__ __align__(16) double rows_sh[kPassWarps][32 * kRowStride];
    PassCtx<MODE> ctx(G, st->T);
    const float cum = st->cum;
    const float cum_eps = __fadd_ru(cum, pp.pos_err);
    unsigned char *my_hard = hard_ids + ((int64_t)blockIdx.x * kPassWarps + warp) * (32 * pp.pts);
    int nhard = 0;  // warp-uniform
    const bool all_inside = st->all_inside != 0;
#pragma unroll 1
    for (int k = 0; k < pp.pts; k++) {
        const int local = (warp * pp.pts + k) * 32 + lane;
        const bool valid = local < task.count;
        bool hard = false, staged = false;
        if (valid) {
            double vs[3];
            ctx.transform(src_xyz, pp.src_n, task.src_begin + local, vs);
            const int slot = task.corr_begin + local;
            if (!all_inside && !inside_grid(G.p, vs[0], vs[1], vs[2])) {
                hot[slot].c0 = -1;  // farther than a cell outside the grid: no neighbour within the radius
            } else {
                hard = true;
                const HotRec h = hot[slot];
                int match = -1;
                double vt[3] = {0.0, 0.0, 0.0}, d2 = 0.0;
                if (pp.use_cache && h.c0 >= 0) {
                    // |q - c0| (exact, rounded up) + move since the proof (upper bound) < distance to anything else
                    // then (lower bound); a NaN or zero limit (knows nothing) compares false
                    const Rec3 t = ld_rec(G.xyz + kPtStride * (int64_t)h.c0);
                    vt[0] = t.x; vt[1] = t.y; vt[2] = t.z;
                    d2 = l2_exact(vs[0], vs[1], vs[2], t.x, t.y, t.z);
                    if (within_limit(__double2float_ru(d2), cum_eps, h.lim1)) {
                        match = h.c0;
                    } else {
                        match = settle_from_set(G, pp, vs[0], vs[1], vs[2], h.c0, slot, cum, cum_eps, hot, cold);
                        if (match >= 0 && match != h.c0) {  // another member of the set won
                            const Rec3 u = ld_rec(G.xyz + kPtStride * (int64_t)match);
                            vt[0] = u.x; vt[1] = u.y; vt[2] = u.z;
                            d2 = l2_exact(vs[0], vs[1], vs[2], u.x, u.y, u.z);
                        }
                    }
                }
                if (match >= 0) {
                    hard = false;
                    // still the nearest, but it may have left the radius: then there is no correspondence
                    // (and nothing to remember: c0 doubles as the correspondence list)
                    if (d2 < pp.r2) {
                        double nt[3] = {0.0, 0.0, 0.0};
                        if (MODE == 1) {
                            const Rec3 nn = ld_rec(G.nrm + kPtStride * (int64_t)match);
                            nt[0] = nn.x; nt[1] = nn.y; nt[2] = nn.z;
                        }
                        ctx.stage_row(rows_sh[warp], vs, vt, nt, d2);
                        staged = true;
                    } else {
                        hot[slot].c0 = -1;
                    }
                }
            }
        }
        if (!staged) ctx.stage_zero_row(rows_sh[warp]);
        VB_STAT(12, hard);
        VB_STAT(13, valid && !hard);
        const unsigned hm = __ballot_sync(0xffffffffu, hard);
        if (hard) my_hard[nhard + __popc(hm & ((1u << lane) - 1u))] = (unsigned char)(k * 32 + lane);
        nhard += __popc(hm);
        if (hm != 0xffffffffu) ctx.accumulate_staged(rows_sh[warp]);  // warp-uniform; nothing to add when every lane is hard
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
    double2 *out = reinterpret_cast<double2 *>(partials + (int64_t)blockIdx.x * kPart);
    if (warp == 0) {
        double2 t = red[lane];
#pragma unroll
        for (int w = 1; w < kPassWarps; w++) { t.x += red[w * 32 + lane].x; t.y += red[w * 32 + lane].y; }
        out[lane] = t;
    }
    // the listed points, as 32-point batches, are part B's work items (one atomic per block that listed anything)
    int total = 0;
#pragma unroll
    for (int w = 0; w < kPassWarps; w++) total += listed[w];
    if (total == 0) return;  // (block-uniform) part B never visits this block: its row is final
    const int nb = (total + 31) >> 5;
    __shared__ int base_sh;
    if (threadIdx.x == 0) base_sh = atomicAdd(pp.work_ctr + 2 * pp.parity, nb);
    __syncthreads();
    if (threadIdx.x < nb) pp.work[base_sh + threadIdx.x] = (blockIdx.x << 4) | threadIdx.x;
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

// ---- pass, part B: the listed points, 32 searches at a time with every lane busy, over a worklist.  Part A
// appends every listed block's batches — 32 of its listed points each — to the list (one atomic per such block);
// this is a resident grid whose WARPS each take batches off the list one at a time (an atomic per batch: dynamic
// balance whatever the number of objects on this GPU, and a near-empty launch once an alignment has settled).
// Which warp handles a batch is arbitrary; what it computes and where it writes — the batch's own Gram row, the
// points' own records — is not.  The last batch of a block to finish folds the block's rows into the block's
// partial row in batch order: row = part A's + batch 0 + batch 1 + ..., the same association whatever ran where,
// so the sums of an object do not depend on what else shares the GPU (the property sharding relies on), and
// k_solve reads ONE row per block.  Each search refreshes the point's records.
template <int MODE>
__global__ void __launch_bounds__(kPassBTpb, kPassBBlocksPerSM) k_pass_b_wl(
    GridDev G, const double *__restrict__ src_xyz, const BlockTask *__restrict__ tasks,
    const ProbState *__restrict__ states, double *partials, HotRec *__restrict__ hot, ColdRec *__restrict__ cold,
    const unsigned char *__restrict__ hard_ids, const int *__restrict__ hard_cnt, PassParams pp) {
    const int lane = threadIdx.x & 31;
    extern __shared__ __align__(16) unsigned char pass_b_smem[];  // kPassBWarps x WarpScratch
    WarpScratch &ws = reinterpret_cast<WarpScratch *>(pass_b_smem)[threadIdx.x >> 5];
    pdl_launch_dependents();
    pdl_wait();  // part A has finished: the worklist and its length are final
    int *ctr = pp.work_ctr + 2 * pp.parity;
    const int n_items = ctr[0];
    if (n_items == 0) return;  // nothing was listed: not even the hand-out counter is touched
#pragma unroll 1
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(ctr + 1, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        const int entry = pp.work[item];
        const int blk = entry >> 4, k = entry & 15;
        int seg_end[kPassWarps];
        int total = 0;
#pragma unroll
        for (int w = 0; w < kPassWarps; w++) {
            total += hard_cnt[blk * kPassWarps + w];
            seg_end[w] = total;
        }
        const BlockTask task = tasks[blk];
        const ProbState *st = states + task.prob;  // (a finished problem's blocks are never listed)
        PassCtx<MODE> ctx(G, st->T);
        const float cum = st->cum;
        // A problem whose latest update moved it by little is settling: its searches look further than the answer
        // needs (as far as everything the problem is still expected to move) and keep candidate sets, so that the
        // next small moves are settled by part A.  One that has just jumped will fail any such test anyway.
        const float last_move = st->last_move;
        const bool settling = last_move < pp.set_move;
        const float slack = settling ? fminf(fmaxf(last_move * (VB_SLACK_MOVE_PCT * 0.01f), pp.slack_lo), pp.slack_hi) : 0.0f;
        {
            const int h = 32 * k + lane;
            const bool live = h < total;
            double x[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            double vs[3] = {0, 0, 0}, d2 = 0.0;
            QueryCtx c;
            int prior = -1, slot = 0;
            if (live) {
                int w = 0, base = 0;
#pragma unroll
                for (int i = 0; i < kPassWarps - 1; i++)
                    if (h >= seg_end[i]) { w = i + 1; base = seg_end[i]; }
                const int id = hard_ids[((int64_t)blk * kPassWarps + w) * (32 * pp.pts) + (h - base)];
                const int local = (w * pp.pts + (id >> 5)) * 32 + (id & 31);
                slot = task.corr_begin + local;
                ctx.transform(src_xyz, pp.src_n, task.src_begin + local, vs);
                make_query(G.p, vs[0], vs[1], vs[2], c);  // inside the grid, or it would not be listed
                prior = hot[slot].c0;  // last iteration's match: a bound for this search
            }
            SearchProof proof;
            const int bs = nn_search_hybrid<32>(G, live, c, vs[0], vs[1], vs[2], pp.r2, pp.r2_ub, prior, ws.runs, &d2,
                                                slack, &proof, VB_COOP_WARP_LANES, settling);
            if (live) {
                HotRec hn;
                hn.c0 = bs;  // sorted position; vb200_batch_corr maps it to the caller's index
                hn.lim1 = bs >= 0 ? lim_of(proof.sec1, cum) : 0.0f;
                hot[slot] = hn;
                ColdRec kn;
                kn.c1 = proof.others[0]; kn.c2 = proof.others[1]; kn.c3 = proof.others[2];
                kn.limK = bs >= 0 ? lim_of(proof.secK, cum) : 0.0f;
                cold[slot] = kn;
                if (bs >= 0) ctx.row_of(bs, vs, pp.r2, x);
            }
            ctx.accumulate(ws.rows, x);
        }
        // the batch's row; then, if this was the block's last batch to finish, fold the block's rows in batch order
        const int nb = (total + 31) >> 5;
        if (nb == 1) {
            // the block's only batch (every listed block of a settled pass): row = part A's + batch 0, as below, without
            // the trip through the batch rows and the block's counter
            double2 *arow = reinterpret_cast<double2 *>(partials + (int64_t)blk * kPart);
            const double2 mine = ctx.lane_pair();
            double2 t = arow[lane];
            t.x += mine.x; t.y += mine.y;
            arow[lane] = t;
            __syncwarp();
            continue;
        }
        double2 *brow = reinterpret_cast<double2 *>(pp.brows + ((int64_t)blk * (kPassWarps * pp.pts)) * kPart);
        brow[k * (kPart / 2) + lane] = ctx.lane_pair();
        fence_release_gpu();  // (not __threadfence(): that would also invalidate the SM's whole L1, once per batch)
        __syncwarp();
        int prev = 0;
        if (lane == 0) prev = atomicAdd(pp.blk_done + blk, 1);
        prev = __shfl_sync(0xffffffffu, prev, 0);
        if (prev == nb - 1) {
            fence_acquire_gpu();
            if (lane == 0) pp.blk_done[blk] = 0;  // ready for the next pass
            double2 *arow = reinterpret_cast<double2 *>(partials + (int64_t)blk * kPart);
            double2 t = arow[lane];  // part A's row (the previous kernel's)
            double2 v[kBatchesPerBlock];
#pragma unroll
            for (int j = 0; j < kBatchesPerBlock; j++)
                v[j] = j < nb ? __ldcg(brow + j * (kPart / 2) + lane) : make_double2(0.0, 0.0);  // written by other SMs
#pragma unroll
            for (int j = 0; j < kBatchesPerBlock; j++)
                if (j < nb) { t.x += v[j].x; t.y += v[j].y; }
            arow[lane] = t;
        }
        __syncwarp();
    }
}

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

// Upper bound of |U q - q| over the points q of a problem whose cloud has bounding sphere (ctr, rad) in its own
// frame and current transform T: |(R_u - I)(q - c) + (R_u - I) c + t_u| <= ||R_u - I||_F * smax(R_T) * rad + |...c...|
// with c = T ctr.  smax(R_T) <= sqrt(max row sum of |R_T^T R_T|) (1 for a rotation; init may be anything).
__device__ double update_move_bound(const double *U, const double *T, const double *ctr, double rad) {
    const double c[3] = {T[0] * ctr[0] + T[1] * ctr[1] + T[2] * ctr[2] + T[3],
                         T[4] * ctr[0] + T[5] * ctr[1] + T[6] * ctr[2] + T[7],
                         T[8] * ctr[0] + T[9] * ctr[1] + T[10] * ctr[2] + T[11]};
    double fro = 0.0, dc2 = 0.0, s2 = 0.0;
    for (int r = 0; r < 3; r++) {
        double dc = U[4 * r + 3];
        for (int k = 0; k < 3; k++) {
            const double m = U[4 * r + k] - (r == k ? 1.0 : 0.0);
            fro += m * m;
            dc += m * c[k];
        }
        dc2 += dc * dc;
        double row = 0.0;  // row r of |R^T R|
        for (int k = 0; k < 3; k++) row += fabs(T[r] * T[k] + T[4 + r] * T[4 + k] + T[8 + r] * T[8 + k]);
        s2 = fmax(s2, row);
    }
    const double d = (sqrt(dc2) + sqrt(fro) * sqrt(s2) * rad) * (1.0 + 1e-9) + 1e-12;
    return d == d ? d : 1e30;  // a NaN update (never produced by the estimators) invalidates every cached bound
}

// result bookkeeping (Registration.cpp:87-93), convergence test (:179-183) and estimator update (:172-174)
// from the reduced slots; one thread.  npts = source points the totals were accumulated over.
__device__ void solve_from_totals(const double *tot, double npts, const ProbDesc &pd, ProbState *st,
                                  const SolveParams &sp, int pass_index) {
    // everything read from the state before anything is written to it
    const double prev_fitness = st->prev_fitness, prev_rmse = st->prev_rmse;
    const float cum = st->cum;
    double U[16], T[16];
    for (int i = 0; i < 16; i++) T[i] = st->T[i];
    const double K = tot[kSlotCount];
    double fitness = 0.0, rmse = 0.0;
    if (K > 0.0 && npts > 0.0) {
        fitness = K / npts;
        rmse = sqrt(tot[kSlotD2] / K);
    }
    st->fitness = fitness;
    st->rmse = rmse;
    st->ncorr = (int)K;
    if (pass_index >= 1 && fabs(prev_fitness - fitness) < sp.rel_fitness && fabs(prev_rmse - rmse) < sp.rel_rmse) {
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
    if (sp.estimator == VB200_EST_P2P) {
        const double cref[3] = {T[3], T[7], T[11]};
        update_p2p(tot, cref, U);
    } else {
        update_p2plane(tot, sp, U);
    }
    const float moved = __double2float_ru(update_move_bound(U, T, pd.ctr, pd.rad));
    mat4_mul(U, T, T);
    // does the cloud's bounding sphere, under the new transform, lie inside the scene grid?  (centre T ctr, radius
    // rad x the largest stretch of T's linear part, bounded by sqrt(max row sum of |R^T R|); margins for the rounding)
    int all_inside = 1;
    {
        double s2 = 0.0;
        for (int r = 0; r < 3; r++) {
            double row = 0.0;
            for (int k = 0; k < 3; k++) row += fabs(T[r] * T[k] + T[4 + r] * T[4 + k] + T[8 + r] * T[8 + k]);
            s2 = fmax(s2, row);
        }
        const double rr = pd.rad * sqrt(s2) * (1.0 + 1e-9);
        for (int a = 0; a < 3; a++) {
            const double c = T[4 * a] * pd.ctr[0] + T[4 * a + 1] * pd.ctr[1] + T[4 * a + 2] * pd.ctr[2] + T[4 * a + 3];
            const double m = 1e-9 * (fabs(c) + rr + 1.0);
            if (!(c - rr - m >= sp.box_lo[a] && c + rr + m <= sp.box_hi[a])) all_inside = 0;  // (NaN: not inside)
        }
    }
    st->last_move = moved;
    st->all_inside = all_inside;
    st->cum = __fadd_ru(cum, moved);
    for (int i = 0; i < 16; i++) st->T[i] = T[i];
    st->iters = pass_index + 1;
}

// Fixed-order reduction of rows [r0, r1) of the partial Gram matrices (kPart doubles each) into sw[0][0..63];
// 256 threads = 4 groups of 64, group g takes rows r0 + g, r0 + g + 4, ... in index order, the groups are then
// added in group order.  kSolveInflight independent loads in flight per thread; the chain of adds keeps its fixed
// order whatever the batch size.
constexpr int kSolveInflight = 16;
__device__ __forceinline__ void reduce_rows(const double *__restrict__ rows, int r0, int r1, double (*sw)[kPart]) {
    const int e = threadIdx.x & (kPart - 1), grp = threadIdx.x / kPart;
    double s = 0.0;
    const double *col = rows + e;
    // (the ragged last batch is predicated, not a loop of dependent loads: adding +0.0 leaves the sum's bits alone)
    for (int b = r0 + grp; b < r1; b += 4 * kSolveInflight) {
        double v[kSolveInflight];
#pragma unroll
        for (int i = 0; i < kSolveInflight; i++) v[i] = b + 4 * i < r1 ? __ldcg(col + (int64_t)(b + 4 * i) * kPart) : 0.0;
#pragma unroll
        for (int i = 0; i < kSolveInflight; i++) s += v[i];
    }
    sw[grp][e] = s;
    __syncthreads();
    if (threadIdx.x < kPart) sw[0][e] = ((sw[0][e] + sw[1][e]) + sw[2][e]) + sw[3][e];
    __syncthreads();
}

// One CLUSTER of kSolveCtas blocks per problem: each block sums a contiguous eighth of the problem's partial rows
// (one round trip to L2 with every load in flight instead of a 32-block kernel walking ~500 rows each), block 0
// adds the eight sums in rank order through distributed shared memory and takes the estimator step.  The order
// of every addition is fixed by the problem's own row count: same bits run to run, alone or in a batch.
constexpr int kSolveCtas = 8;
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ld_dsmem_f64(const double *local_smem_ptr, unsigned rank) {
    unsigned a = (unsigned)__cvta_generic_to_shared(local_smem_ptr), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
    return v;
}

// the cluster's reduction: afterwards block 0's tot[0..31] holds the problem's estimator slots
__device__ __forceinline__ void cluster_reduce_partials(const ProbDesc &pd, const double *__restrict__ partials,
                                                        bool plane, double (*sw)[kPart], double *tot) {
    const unsigned rank = cluster_ctarank();
    const int nrows = pd.blk_count;  // one row per block (part B folds its batches' rows into it)
    const int per = (nrows + kSolveCtas - 1) / kSolveCtas;
    const int r0 = min((int)rank * per, nrows), r1 = min(r0 + per, nrows);
    reduce_rows(partials + (int64_t)pd.blk_begin * kPart, r0, r1, sw);
    cluster_sync_all();  // every block's sw[0] is final and visible cluster-wide
    if (rank == 0) {
        if (threadIdx.x < kPart) {
            double t = sw[0][threadIdx.x];
#pragma unroll
            for (unsigned r = 1; r < kSolveCtas; r++) t += ld_dsmem_f64(&sw[0][threadIdx.x], r);
            sw[1][threadIdx.x] = t;
        }
        __syncthreads();
        if (threadIdx.x < kAcc) tot[threadIdx.x] = slot_from_gram(sw[1], plane, threadIdx.x);
        __syncthreads();
    }
    cluster_sync_all();  // nobody exits while block 0 may still read its shared memory
}

__global__ void __launch_bounds__(256) k_solve(const ProbDesc *__restrict__ probs, ProbState *__restrict__ states,
                                               const double *__restrict__ partials, SolveParams sp, int pass_index) {
    const int p = blockIdx.x / kSolveCtas;
    const ProbDesc pd = probs[p];  // static since set_problems
    pdl_launch_dependents();
    pdl_wait();  // both parts of the pass have written their partial rows
    ProbState *st = states + p;
    if (st->done) return;  // (cluster-uniform)
    __shared__ double sw[4][kPart];
    __shared__ double tot[kAcc];
    cluster_reduce_partials(pd, partials, sp.estimator != VB200_EST_P2P, sw, tot);
    if (cluster_ctarank() == 0 && threadIdx.x == 0) solve_from_totals(tot, (double)pd.npts, pd, st, sp, pass_index);
}

// The same tail split in two for multi-GPU alignment of ONE cloud sharded over ranks (ICPRefinement's global
// transform, src/evaluation.cpp:244-274): k_reduce leaves each problem's 32 totals in a device buffer the
// caller all-reduces across GPUs, k_solve_totals finishes the iteration from the combined totals.  Every rank
// sees identical totals, hence applies the identical update: no broadcast of T is needed.
__global__ void __launch_bounds__(256) k_reduce(const ProbDesc *__restrict__ probs, const ProbState *__restrict__ states,
                                                const double *__restrict__ partials, bool plane,
                                                double *__restrict__ totals) {
    const int p = blockIdx.x / kSolveCtas;
    const ProbDesc pd = probs[p];
    pdl_launch_dependents();
    pdl_wait();
    __shared__ double sw[4][kPart];
    __shared__ double tot[kAcc];
    if (states[p].done) return;  // finished problems contribute their last totals unchanged
    cluster_reduce_partials(pd, partials, plane, sw, tot);
    if (cluster_ctarank() == 0 && threadIdx.x < kAcc) totals[(int64_t)p * kAcc + threadIdx.x] = tot[threadIdx.x];
}

__global__ void __launch_bounds__(32) k_solve_totals(int P, const ProbDesc *__restrict__ probs,
                                                     ProbState *__restrict__ states,
                                                     const double *__restrict__ totals,
                                                     const double *__restrict__ npts_global, SolveParams sp,
                                                     int pass_index) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P || states[p].done) return;
    double tot[kAcc];
    for (int i = 0; i < kAcc; i++) tot[i] = totals[(int64_t)p * kAcc + i];
    solve_from_totals(tot, npts_global[p], probs[p], states + p, sp, pass_index);
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
    // structure of arrays (x[n], y[n], z[n]): a warp's 32 consecutive points are three 256-byte requests
    // (8 sectors each) instead of three strided ones of 24 sectors
    out[s] = in[3 * (int64_t)i];
    out[(int64_t)n + s] = in[3 * (int64_t)i + 1];
    out[2 * (int64_t)n + s] = in[3 * (int64_t)i + 2];
    orig[s] = i - cloud_off[find_cloud(cloud_off, ncloud, i)];
}

// correspondences are kept as sorted scene positions (they double as the next pass's search bound); this maps
// one problem's column to the caller's target indices
__global__ void __launch_bounds__(256) k_corr_orig(const HotRec *__restrict__ hot, int n, const int *__restrict__ orig,
                                                   int *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = hot[i].c0;
    out[i] = s >= 0 ? orig[s] : -1;
}

// ---- per-cloud bounding sphere (what update_move_bound needs), computed once at upload ------------------
__device__ __forceinline__ unsigned long long ordered_bits(double v) {  // monotone map double -> uint64
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ inline double from_ordered_bits(unsigned long long o) {
    const unsigned long long u = (o >> 63) ? (o & 0x7fffffffffffffffull) : ~o;
    double v;
    memcpy(&v, &u, sizeof(v));
    return v;
}

__global__ void __launch_bounds__(256) k_cloud_bounds_init(unsigned long long *__restrict__ bounds, int ncloud) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 7 * ncloud) bounds[i] = (i < 6 * ncloud && i % 6 < 3) ? ~0ull : 0ull;
}

// bounds[c] = {min x, y, z, max x, y, z} as ordered bits (initialised to ~0 / 0 by the caller).  A warp whose
// lanes all belong to one cloud (nearly every warp) reduces with shuffles and issues six atomics.
__global__ void __launch_bounds__(256) k_cloud_bbox(const double *__restrict__ xyz, int n,
                                                    const int *__restrict__ cloud_off, int ncloud,
                                                    unsigned long long *__restrict__ bounds) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const int c = valid ? find_cloud(cloud_off, ncloud, i) : -1;
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    if (!act) return;
    const int c_first = __shfl_sync(0xffffffffu, c, __ffs(act) - 1);
    const bool uniform = __all_sync(0xffffffffu, !valid || c == c_first);
    for (int a = 0; a < 3; a++) {
        unsigned long long lo = valid ? ordered_bits(xyz[3 * (int64_t)i + a]) : ~0ull, hi = valid ? lo : 0ull;
        if (uniform) {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
                lo = l2 < lo ? l2 : lo;
                hi = h2 > hi ? h2 : hi;
            }
            if ((threadIdx.x & 31) == 0) {
                atomicMin(bounds + 6 * c_first + a, lo);
                atomicMax(bounds + 6 * c_first + 3 + a, hi);
            }
        } else if (valid) {
            atomicMin(bounds + 6 * c + a, lo);
            atomicMax(bounds + 6 * c + 3 + a, hi);
        }
    }
}

// rad2[c] = max |p - centre|^2 (bits of a non-negative double order like integers)
__global__ void __launch_bounds__(256) k_cloud_radius(const double *__restrict__ xyz, int n,
                                                      const int *__restrict__ cloud_off, int ncloud,
                                                      const unsigned long long *__restrict__ bounds,
                                                      unsigned long long *__restrict__ rad2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const int c = valid ? find_cloud(cloud_off, ncloud, i) : -1;
    double d2 = 0.0;
    if (valid) {
        for (int a = 0; a < 3; a++) {
            const double ctr = 0.5 * (from_ordered_bits(bounds[6 * c + a]) + from_ordered_bits(bounds[6 * c + 3 + a]));
            const double d = xyz[3 * (int64_t)i + a] - ctr;
            d2 += d * d;
        }
        if (!(d2 == d2)) d2 = 0.0;
    }
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    if (!act) return;
    const int c_first = __shfl_sync(0xffffffffu, c, __ffs(act) - 1);
    unsigned long long m = (unsigned long long)__double_as_longlong(d2);
    if (__all_sync(0xffffffffu, !valid || c == c_first)) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const unsigned long long m2 = __shfl_xor_sync(0xffffffffu, m, o);
            m = m2 > m ? m2 : m;
        }
        if ((threadIdx.x & 31) == 0) atomicMax(rad2 + c_first, m);
    } else if (valid) {
        atomicMax(rad2 + c, m);
    }
}

// ---- estimator plug-in kernels: reductions over an explicit correspondence list ------------------------
// corr == nullptr: row i of vs_in / vt_in / nt_in is correspondence i (the host entry gathers them, as the
// reference does, TransformationEstimation.cpp:52-57); otherwise corr[2i], corr[2i+1] index the clouds in place
// (device-resident entry: nothing is gathered on the host).
template <int MODE>
__global__ void __launch_bounds__(kPassTpb) k_estimate(const double *__restrict__ vs_in,
                                                       const double *__restrict__ vt_in,
                                                       const double *__restrict__ nt_in,
                                                       const int *__restrict__ corr, int64_t K,
                                                       double *__restrict__ partials) {
    __shared__ double swarp[kPassTpb / 32][kAcc];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t first = corr ? corr[0] : 0;
    const double cref[3] = {vs_in[3 * first], vs_in[3 * first + 1], vs_in[3 * first + 2]};
    double acc = 0.0;
#pragma unroll 1
    for (int k = 0; k < kPtsPerThread; k++) {
        int64_t i = (int64_t)blockIdx.x * kChunk + k * kPassTpb + threadIdx.x;
        bool valid = i < K;
        double vs[3] = {0, 0, 0}, vt[3] = {0, 0, 0}, nt[3] = {0, 0, 0};
        if (valid) {
            const int64_t is = corr ? corr[2 * i] : i, it = corr ? corr[2 * i + 1] : i;
            for (int a = 0; a < 3; a++) {
                vs[a] = vs_in[3 * is + a];
                vt[a] = vt_in[3 * it + a];
                if (MODE == 1) nt[a] = nt_in[3 * it + a];
            }
        }
        double v[kAcc];
        contributions<MODE>(valid, vs, vt, nt, cref, v);
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

// out[0..15] = the estimator's transformation; out[16] = sqrt(sum |s - t|^2 / K), the point-to-point
// ComputeRMSE (src/constrained_ICP.cpp:13-23, TransformationEstimation.cpp:35-45)
__global__ void __launch_bounds__(256) k_estimate_solve(const double *__restrict__ partials, int nblk,
                                                        const double *__restrict__ vs_in,
                                                        const int *__restrict__ corr, SolveParams sp,
                                                        double *__restrict__ out) {
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
        const int64_t first = corr ? corr[0] : 0;
        const double cref[3] = {vs_in[3 * first], vs_in[3 * first + 1], vs_in[3 * first + 2]};
        update_p2p(tot, cref, U);
    } else {
        update_p2plane(tot, sp, U);
    }
    for (int i = 0; i < 16; i++) out[i] = U[i];
    out[16] = tot[kSlotCount] > 0.0 ? sqrt(tot[kSlotRes2] / tot[kSlotCount]) : 0.0;
}

}  // namespace

// ======================================================================================================
struct Batch {
    Scene *scene = nullptr;
    cudaStream_t stream = nullptr;  // the scene's stream, or its second one (vb200_icp_run pipelines two halves)
    int ncloud = 0;
    int64_t npts = 0;
    std::vector<int> cloud_off;     // host copy, ncloud+1
    std::vector<double> cloud_sphere;  // per cloud: centre of its bounding box (3) and bounding radius about it
    bool has_normals = false;
    bool use_cache = true;          // vb200_batch_set_option(VB200_OPT_NN_CACHE)
    bool split_timing = false;      // vb200_batch_set_option(VB200_OPT_SPLIT_TIMING)
    // programmatic dependent launch between the kernels of an iteration.  Off when two batches run concurrently on
    // two streams (vb200_icp_run's halves): a dependent grid that is resident early and waiting holds the registers
    // the OTHER stream's kernels need (measured: 4.6 vs 4.1 ms per converging 32-object call)
    bool chain = true;
    double *d_src = nullptr;        // sorted source points as x[npts], y[npts], z[npts]
    int *d_src_orig = nullptr;      // sorted position -> original index local to its cloud
    int *d_cloud_off = nullptr;
    // problems
    int P = 0;
    std::vector<ProbDesc> probs;
    int nblk = 0;
    int64_t ncorr_slots = 0;
    ProbDesc *d_probs = nullptr;
    ProbState *d_states = nullptr;
    BlockTask *d_tasks = nullptr;
    double *d_partials = nullptr;
    HotRec *d_hot = nullptr;                 // per (problem, point): match + single-candidate bound (k_pass_a/b)
    ColdRec *d_cold = nullptr;               // ... and the rest of its candidate set
    unsigned char *d_hard_ids = nullptr;     // per block and warp: the points part A left for part B
    int *d_hard_cnt = nullptr;
    int *d_ndone = nullptr;                  // problems finished since set_problems
    int *d_work = nullptr;                   // part B's work items: (block << 4 | batch)  (k_pass_b_wl)
    int *d_work_ctr = nullptr;               // 2 x {listed, handed out}
    double *d_brows = nullptr;               // per (block, batch): the batch's Gram row until the block's rows are folded
    int *d_blk_done = nullptr;               // per block: batches finished in the current pass
    int pass_parity = 0;
    // Points per part-A thread (4, 2 or 1; a block covers 128 x pts points).  A warp walks its pts batches one after the
    // other, three dependent round trips each: with many objects on the GPU four batches per warp keep the block count
    // (and the per-block partial rows) down at no cost, with few — 4 objects per GPU when 32 are sharded over 8 — the
    // blocks would not fill the machine and the walk's latency chain is the whole kernel (ncu: 20 us for 11 MB), so the
    // points are spread over more, shorter-lived blocks.
    int pts = kPtsPerThread;
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
    void *ptrs[] = {b->d_probs, b->d_states, b->d_tasks, b->d_partials, b->d_hot, b->d_cold, b->d_totals,
                    b->d_npts_global, b->d_hard_ids, b->d_hard_cnt, b->d_ndone, b->d_work, b->d_work_ctr,
                    b->d_brows, b->d_blk_done};
    for (void *q : ptrs)
        if (q) cudaFreeAsync(q, st);
    b->d_probs = nullptr; b->d_states = nullptr; b->d_tasks = nullptr; b->d_partials = nullptr;
    b->d_hot = nullptr; b->d_cold = nullptr; b->d_totals = nullptr; b->d_npts_global = nullptr;
    b->d_hard_ids = nullptr; b->d_hard_cnt = nullptr; b->d_ndone = nullptr; b->d_work = nullptr;
    b->d_work_ctr = nullptr; b->d_brows = nullptr; b->d_blk_done = nullptr;
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
    b->cloud_sphere.assign(4 * (size_t)std::max(ncloud, 1), 0.0);
    VB_CUDA(cudaMallocAsync((void **)&b->d_cloud_off, sizeof(int) * ((size_t)ncloud + 1), st));
    VB_CUDA(cudaMemcpyAsync(b->d_cloud_off, b->cloud_off.data(), sizeof(int) * ((size_t)ncloud + 1),
                            cudaMemcpyHostToDevice, st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_src, sizeof(double) * 3 * (size_t)std::max(n, 1), st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_src_orig, sizeof(int) * (size_t)std::max(n, 1), st));
    if (n == 0) return VB200_OK;
    DevBuf<double> d_in(st);
    DevBuf<int> d_key(st), d_counts(st), d_start(st), d_sidx(st);
    DevBuf<unsigned long long> d_bounds(st);  // per cloud: 6 bbox words + 1 radius word
    const size_t nb = (size_t)ncloud * kBuckets + 1;
    VB_CUDA(d_in.alloc(3 * (size_t)n));
    VB_CUDA(d_key.alloc((size_t)n));
    VB_CUDA(d_sidx.alloc((size_t)n));
    VB_CUDA(d_counts.alloc(nb));
    VB_CUDA(d_start.alloc(nb));
    VB_CUDA(d_bounds.alloc(7 * (size_t)ncloud));
    VB_CUDA(h2d_async(d_in.p, src_xyz + 3 * off[0], sizeof(double) * 3 * (size_t)n, st));
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
    // bounding sphere of every cloud: min words start at all-ones, max and radius words at zero
    k_cloud_bounds_init<<<div_up(7 * (int64_t)ncloud, 256), 256, 0, st>>>(d_bounds.p, ncloud);
    k_cloud_bbox<<<div_up(n, 256), 256, 0, st>>>(d_in.p, n, b->d_cloud_off, ncloud, d_bounds.p);
    k_cloud_radius<<<div_up(n, 256), 256, 0, st>>>(d_in.p, n, b->d_cloud_off, ncloud, d_bounds.p,
                                                   d_bounds.p + 6 * (size_t)ncloud);
    VB_CUDA(cudaGetLastError());
    std::vector<unsigned long long> hb(7 * (size_t)ncloud);
    VB_CUDA(cudaMemcpyAsync(hb.data(), d_bounds.p, sizeof(unsigned long long) * hb.size(), cudaMemcpyDeviceToHost, st));
    b->launches += 5 + 3 + 2;
    VB_CUDA(cudaStreamSynchronize(st));  // temporaries are released on return
    for (int c = 0; c < ncloud; c++) {
        if (b->cloud_off[c + 1] == b->cloud_off[c]) continue;
        double r2;
        memcpy(&r2, &hb[6 * (size_t)ncloud + c], sizeof(r2));
        for (int a = 0; a < 3; a++)
            b->cloud_sphere[4 * (size_t)c + a] = 0.5 * (from_ordered_bits(hb[6 * (size_t)c + a]) + from_ordered_bits(hb[6 * (size_t)c + 3 + a]));
        b->cloud_sphere[4 * (size_t)c + 3] = sqrt(r2) * (1.0 + 1e-12);
    }
    return VB200_OK;
}

static int batch_set_problems_impl(Batch *b, const int32_t *cloud_ids, const double *init_T, int P) {
    cudaStream_t st = b->stream;
    // validate everything before touching the batch: a bad argument leaves the previous problems intact
    {
        int64_t total = 0;
        for (int p = 0; p < P; p++) {
            const int c = cloud_ids ? cloud_ids[p] : p;
            if (c < 0 || c >= b->ncloud) return VB200_ERR_INVALID;
            total += b->cloud_off[c + 1] - b->cloud_off[c];
            if (total > 0x7fffffff) return VB200_ERR_INVALID;
        }
    }
    batch_free_problems(b);  // P = 0 from here until every upload below has succeeded
    b->iter_base = 0;
    // points per part-A thread: the largest of 4, 2, 1 that still gives part A one full wave of blocks
    b->pts = kPtsPerThread;
    for (;;) {
        int64_t nb = 0;
        for (int p = 0; p < P; p++) {
            const int c = cloud_ids ? cloud_ids[p] : p;
            nb += div_up(b->cloud_off[c + 1] - b->cloud_off[c], kPassTpb * b->pts);
        }
        if (b->pts == 1 || nb >= (int64_t)kNumSMsB200 * VB_PASS_A_MINBLOCKS) break;
        b->pts >>= 1;
    }
    const int chunk = kPassTpb * b->pts, batches_per_block = kPassWarps * b->pts;
    b->probs.resize((size_t)P);
    std::vector<BlockTask> tasks;
    std::vector<ProbState> states((size_t)P);
    int64_t corr = 0;
    for (int p = 0; p < P; p++) {
        int c = cloud_ids ? cloud_ids[p] : p;
        ProbDesc &pd = b->probs[p];
        pd.cloud = c;
        pd.npts = b->cloud_off[c + 1] - b->cloud_off[c];
        pd.src_begin = b->cloud_off[c];
        pd.corr_begin = (int)corr;
        pd.blk_begin = (int)tasks.size();
        for (int a = 0; a < 3; a++) pd.ctr[a] = b->cloud_sphere[4 * (size_t)c + a];
        pd.rad = b->cloud_sphere[4 * (size_t)c + 3];
        for (int s = 0; s < pd.npts; s += chunk) {
            BlockTask t;
            t.prob = p;
            t.src_begin = pd.src_begin + s;
            t.count = std::min(chunk, pd.npts - s);
            t.corr_begin = pd.corr_begin + s;
            tasks.push_back(t);
        }
        pd.blk_count = (int)tasks.size() - pd.blk_begin;
        corr += pd.npts;
        ProbState &s = states[p];
        memset(&s, 0, sizeof(s));
        for (int i = 0; i < 16; i++) s.T[i] = init_T[16 * (size_t)p + i];
        s.last_move = 1e30f;  // nothing is "settling" before the first update
    }
    b->nblk = (int)tasks.size();
    b->ncorr_slots = corr;
    const size_t nslot = (size_t)std::max<int64_t>(corr, 1), nblk1 = (size_t)std::max(b->nblk, 1), P1 = (size_t)std::max(P, 1);
    VB_CUDA(cudaMallocAsync((void **)&b->d_probs, sizeof(ProbDesc) * P1, st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_states, sizeof(ProbState) * P1, st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_tasks, sizeof(BlockTask) * nblk1, st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_partials, sizeof(double) * kPart * nblk1, st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_brows, sizeof(double) * kPart * (size_t)batches_per_block * nblk1, st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_blk_done, sizeof(int) * nblk1, st));
    VB_CUDA(cudaMemsetAsync(b->d_blk_done, 0, sizeof(int) * nblk1, st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_hot, sizeof(HotRec) * nslot, st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_cold, sizeof(ColdRec) * nslot, st));
    if (P) {
        VB_CUDA(cudaMemcpyAsync(b->d_probs, b->probs.data(), sizeof(ProbDesc) * (size_t)P, cudaMemcpyHostToDevice, st));
        VB_CUDA(cudaMemcpyAsync(b->d_states, states.data(), sizeof(ProbState) * (size_t)P, cudaMemcpyHostToDevice, st));
    }
    if (b->nblk)
        VB_CUDA(cudaMemcpyAsync(b->d_tasks, tasks.data(), sizeof(BlockTask) * (size_t)b->nblk, cudaMemcpyHostToDevice, st));
    // 0xff: c = -1 (no match, no candidates), limits NaN (every test compares false): knows nothing
    VB_CUDA(cudaMemsetAsync(b->d_hot, 0xff, sizeof(HotRec) * nslot, st));
    VB_CUDA(cudaMemsetAsync(b->d_cold, 0xff, sizeof(ColdRec) * nslot, st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_hard_ids, (size_t)chunk * nblk1, st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_hard_cnt, sizeof(int) * kPassWarps * nblk1, st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_ndone, sizeof(int), st));
    VB_CUDA(cudaMemsetAsync(b->d_ndone, 0, sizeof(int), st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_work, sizeof(int) * (size_t)batches_per_block * nblk1, st));
    VB_CUDA(cudaMallocAsync((void **)&b->d_work_ctr, sizeof(int) * 4, st));
    VB_CUDA(cudaMemsetAsync(b->d_work_ctr, 0, sizeof(int) * 4, st));
    b->pass_parity = 0;
    VB_CUDA(cudaStreamSynchronize(st));  // host vectors go out of scope
    b->P = P;
    return VB200_OK;
}

static int batch_set_problems(Batch *b, const int32_t *cloud_ids, const double *init_T, int P) {
    const int rc = batch_set_problems_impl(b, cloud_ids, init_T, P);
    // a failed set-up (CUDA error half-way) must not leave a batch that later launches on null buffers
    if (rc != VB200_OK && rc != VB200_ERR_INVALID) batch_free_problems(b);
    return rc;
}

// A kernel launch chained to the previous one on the stream by programmatic dependent launch (the kernels call
// pdl_wait() before touching anything the previous kernel wrote), optionally as clusters of `cluster` blocks.
template <typename... KArgs, typename... Args>
static cudaError_t launch_chained(void (*kernel)(KArgs...), int grid, int block, int cluster, size_t smem,
                                  cudaStream_t st, bool chain, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3((unsigned)block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (chain) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        na++;
    }
    if (cluster > 1) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = (unsigned)cluster;
        at[na].val.clusterDim.y = 1;
        at[na].val.clusterDim.z = 1;
        na++;
    }
    cfg.attrs = at;
    cfg.numAttrs = (unsigned)na;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// one correspondence pass = part A (every point, streaming) + part B (the points that need a search)
static int launch_pass(Batch *b, bool plane, const PassParams &pp_in) {
    Scene *sc = b->scene;
    cudaStream_t st = b->stream;
    PassParams pp = pp_in;
    pp.work = b->d_work;
    pp.work_ctr = b->d_work_ctr;
    pp.parity = b->pass_parity;
    pp.brows = b->d_brows;
    pp.blk_done = b->d_blk_done;
    pp.src_n = b->npts;
    pp.pts = b->pts;
    b->pass_parity ^= 1;
    // part B's scratch is dynamic shared memory (more than 48 KB per block when the block is large): opt in once
    // per device and kernel
    constexpr size_t kPassBSmem = sizeof(WarpScratch) * kPassBWarps;
    if (kPassBSmem > 48 * 1024) {
        static bool opted[64][2] = {};
        const int dev = sc->device;
        if (dev >= 64 || !opted[dev][plane ? 1 : 0]) {
            if (plane) VB_CUDA(cudaFuncSetAttribute(k_pass_b_wl<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPassBSmem));
            else VB_CUDA(cudaFuncSetAttribute(k_pass_b_wl<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPassBSmem));
            if (dev < 64) opted[dev][plane ? 1 : 0] = true;
        }
    }
    const int nwarps = (int)std::min<int64_t>((int64_t)b->nblk * kPassWarps * b->pts, kNumSMsB200 * kPassBWarps * kPassBBlocksPerSM);
    const int nb_b = div_up(nwarps, kPassBWarps);
    if (plane) {
        VB_CUDA(launch_chained(k_pass_a<1>, b->nblk, kPassTpb, 1, 0, st, b->chain, sc->grid, (const double *)b->d_src,
                               (const BlockTask *)b->d_tasks, (const ProbState *)b->d_states, b->d_partials, b->d_hot,
                               b->d_cold, b->d_hard_ids, b->d_hard_cnt, pp));
        VB_CUDA(launch_chained(k_pass_b_wl<1>, nb_b, kPassBTpb, 1, kPassBSmem, st, b->chain, sc->grid, (const double *)b->d_src,
                               (const BlockTask *)b->d_tasks, (const ProbState *)b->d_states, b->d_partials, b->d_hot,
                               b->d_cold, (const unsigned char *)b->d_hard_ids, (const int *)b->d_hard_cnt, pp));
    } else {
        VB_CUDA(launch_chained(k_pass_a<0>, b->nblk, kPassTpb, 1, 0, st, b->chain, sc->grid, (const double *)b->d_src,
                               (const BlockTask *)b->d_tasks, (const ProbState *)b->d_states, b->d_partials, b->d_hot,
                               b->d_cold, b->d_hard_ids, b->d_hard_cnt, pp));
        VB_CUDA(launch_chained(k_pass_b_wl<0>, nb_b, kPassBTpb, 1, kPassBSmem, st, b->chain, sc->grid, (const double *)b->d_src,
                               (const BlockTask *)b->d_tasks, (const ProbState *)b->d_states, b->d_partials, b->d_hot,
                               b->d_cold, (const unsigned char *)b->d_hard_ids, (const int *)b->d_hard_cnt, pp));
    }
    b->launches += 2;
    return VB200_OK;
}

static int launch_solve(Batch *b, const SolveParams &sp, int pass_index) {
    VB_CUDA(launch_chained(k_solve, b->P * kSolveCtas, 256, kSolveCtas, 0, b->stream, b->chain, (const ProbDesc *)b->d_probs,
                           b->d_states, (const double *)b->d_partials, sp, pass_index));
    b->launches++;
    return VB200_OK;
}

static int make_params(Batch *b, int estimator, const double *gravity, double max_dist, PassParams *pp,
                       SolveParams *sp) {
    Scene *sc = b->scene;
    if (estimator < VB200_EST_P2P || estimator > VB200_EST_P2PLANE_GRAVITY) return VB200_ERR_INVALID;
    if (!(max_dist > 0.0)) return VB200_ERR_DISTANCE;
    if (max_dist > sc->grid.p.cell * (1.0 + 1e-12)) return VB200_ERR_INVALID;
    if (estimator != VB200_EST_P2P && (!b->has_normals || !sc->has_normals)) return VB200_ERR_NORMALS;
    *pp = make_pass_params(sc->grid.p, max_dist, b->use_cache);
    memset(sp, 0, sizeof(*sp));
    sp->ndone = nullptr;
    sp->estimator = estimator;
    for (int a = 0; a < 3; a++) {
        // inside_grid accepts (x - lo) * inv_fine in [0, fdim): a box shrunk by a relative 1e-9 of its extent on each side
        const GridParams &g = sc->grid.p;
        const double ext = (double)g.fdim[a] / g.inv_fine;
        sp->box_lo[a] = g.lo[a] + 1e-9 * (ext + fabs(g.lo[a]));
        sp->box_hi[a] = g.lo[a] + ext - 1e-9 * (ext + fabs(g.lo[a]));
    }
    sp->g[0] = 0.0; sp->g[1] = 1.0; sp->g[2] = 0.0;  // VISMA's gravity convention: +Y (src/annotation.cpp:43,84)
    if (estimator == VB200_EST_P2PLANE_GRAVITY && gravity) {
        double l = sqrt(gravity[0] * gravity[0] + gravity[1] * gravity[1] + gravity[2] * gravity[2]);
        if (!(l > 0.0)) return VB200_ERR_INVALID;
        for (int a = 0; a < 3; a++) sp->g[a] = gravity[a] / l;
    }
    return VB200_OK;
}

// Iterations [it_begin, it_end] of the loop (the whole loop is 0..max_iter): vb200_icp_run enqueues the first few
// of one half of its clouds, uploads the other half meanwhile, then comes back for the rest.
static int batch_run(Batch *b, int estimator, const double *gravity, double max_dist, double rel_fitness,
                     double rel_rmse, int max_iter, int it_begin = 0, int it_end = 0x7fffffff, bool poll = true) {
    cudaStream_t st = b->stream;
    if (max_iter < 0) return VB200_ERR_INVALID;
    PassParams pp;
    SolveParams sp;
    VB_TRY(make_params(b, estimator, gravity, max_dist, &pp, &sp));
    if (b->P == 0) return VB200_OK;
    const bool plane = estimator != VB200_EST_P2P;
    sp.rel_fitness = rel_fitness;
    sp.rel_rmse = rel_rmse;
    sp.max_iter = max_iter;
    // With a live convergence test most runs finish long before max_iter (5-7 iterations on the BASELINE
    // workload): every fourth iteration the host looks at the finished-problems counter and stops enqueueing
    // passes that would only find every problem done.
    const bool can_stop = rel_fitness > 0.0 && rel_rmse > 0.0;
    if (can_stop) sp.ndone = b->d_ndone;
    for (int it = it_begin; it <= std::min(max_iter, it_end); it++) {
        if (poll && can_stop && it >= 4 && (it & 3) == 0) {
            int ndone = 0;
            VB_CUDA(cudaMemcpyAsync(&ndone, b->d_ndone, sizeof(int), cudaMemcpyDeviceToHost, st));
            VB_CUDA(cudaStreamSynchronize(st));
            if (ndone >= b->P) break;
        }
        if (b->nblk) VB_TRY(launch_pass(b, plane, pp));
        VB_TRY(launch_solve(b, sp, it));
    }
    VB_CUDA(cudaGetLastError());
    return VB200_OK;
}

// has every problem of the batch finished?  (waits for what has been enqueued on the batch's stream)
static int batch_all_done(Batch *b, bool *done) {
    *done = true;
    if (b->P == 0 || !b->d_ndone) return VB200_OK;
    int ndone = 0;
    VB_CUDA(cudaMemcpyAsync(&ndone, b->d_ndone, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    VB_CUDA(cudaStreamSynchronize(b->stream));
    *done = ndone >= b->P;
    return VB200_OK;
}

// first half of an iteration: correspondence pass + per-problem totals left on the device
static int batch_pass(Batch *b, int estimator, double max_dist) {
    cudaStream_t st = b->stream;
    PassParams pp;
    SolveParams sp;
    VB_TRY(make_params(b, estimator, nullptr, max_dist, &pp, &sp));
    if (b->P == 0) return VB200_OK;
    if (!b->d_totals && !b->d_totals_ext) {
        VB_CUDA(cudaMallocAsync((void **)&b->d_totals, sizeof(double) * kAcc * (size_t)b->P, st));
        VB_CUDA(cudaMemsetAsync(b->d_totals, 0, sizeof(double) * kAcc * (size_t)b->P, st));
    }
    double *totals = b->d_totals_ext ? b->d_totals_ext : b->d_totals;
    if (b->nblk) VB_TRY(launch_pass(b, estimator != VB200_EST_P2P, pp));
    VB_CUDA(launch_chained(k_reduce, b->P * kSolveCtas, 256, kSolveCtas, 0, st, b->chain, (const ProbDesc *)b->d_probs,
                           (const ProbState *)b->d_states, (const double *)b->d_partials,
                           estimator != VB200_EST_P2P, totals));
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
    k_solve_totals<<<div_up(b->P, 32), 32, 0, st>>>(b->P, b->d_probs, b->d_states, totals, b->d_npts_global, sp, pass_index);
    b->launches++;
    VB_CUDA(cudaGetLastError());
    VB_CUDA(cudaStreamSynchronize(st));  // `np` is pageable host memory
    return VB200_OK;
}

// n unconditional iterations from the current transforms (no convergence test, `done` never set).  With
// VB200_OPT_SPLIT_TIMING the pass and the solve of a single iteration are bracketed by events (which keeps the
// solve from overlapping the pass's tail: a diagnostic, not the way the loop normally runs).
static int batch_iterate(Batch *b, int estimator, const double *gravity, double max_dist, int n_iter) {
    cudaStream_t st = b->stream;
    if (n_iter < 0) return VB200_ERR_INVALID;
    PassParams pp;
    SolveParams sp;
    VB_TRY(make_params(b, estimator, gravity, max_dist, &pp, &sp));
    if (b->P == 0 || b->nblk == 0) return VB200_OK;
    const bool plane = estimator != VB200_EST_P2P;
    sp.rel_fitness = sp.rel_rmse = -1.0;  // |delta| < -1 never holds: no convergence exit
    sp.max_iter = 0x7fffffff;
    for (int i = 0; i < 3; i++)
        if (!b->ev[i]) VB_CUDA(cudaEventCreate(&b->ev[i]));
    const bool timed = b->split_timing && n_iter == 1;  // per-kernel events only make sense around a single iteration
    for (int it = 0; it < n_iter; it++) {
        if (timed) VB_CUDA(cudaEventRecord(b->ev[0], st));
        VB_TRY(launch_pass(b, plane, pp));
        if (timed) VB_CUDA(cudaEventRecord(b->ev[1], st));
        VB_TRY(launch_solve(b, sp, b->iter_base++));
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
        vb::k_corr_orig<<<vb::div_up(pd.npts, 256), 256, 0, b->stream>>>(b->d_hot + pd.corr_begin, pd.npts,
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
        if (rc == VB200_OK && nhalf == 2) reinterpret_cast<Batch *>(half[h])->chain = false;
        if (rc == VB200_OK) rc = vb200_batch_set_problems(half[h], nullptr, init_T + 16 * (size_t)p0, np);
        if (rc == VB200_OK)
            rc = vb::batch_run(reinterpret_cast<Batch *>(half[h]), estimator, gravity_axis, max_dist, rel_fitness,
                               rel_rmse, max_iter, 0, nhalf == 2 ? warm : 0x7fffffff);
    }
    if (nhalf == 2 && rel_fitness > 0.0 && rel_rmse > 0.0) {
        // a live convergence test: four iterations of BOTH halves are enqueued, then the host looks at both finished-
        // problems counters — the halves' tails overlap on the device instead of the second waiting for the first's polls
        bool fin[2] = {false, false};
        for (int it0 = warm + 1; it0 <= max_iter && rc == VB200_OK && !(fin[0] && fin[1]); it0 += 4) {
            for (int h = 0; h < 2 && rc == VB200_OK; h++)
                if (!fin[h])
                    rc = vb::batch_run(reinterpret_cast<Batch *>(half[h]), estimator, gravity_axis, max_dist, rel_fitness,
                                       rel_rmse, max_iter, it0, it0 + 3, false);
            for (int h = 0; h < 2 && rc == VB200_OK; h++)
                if (!fin[h]) rc = vb::batch_all_done(reinterpret_cast<Batch *>(half[h]), &fin[h]);
        }
    } else {
        for (int h = 0; h < nhalf && rc == VB200_OK && nhalf == 2; h++)
            rc = vb::batch_run(reinterpret_cast<Batch *>(half[h]), estimator, gravity_axis, max_dist, rel_fitness, rel_rmse,
                               max_iter, warm + 1);
    }
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

// shared tail of the estimator plug-in entries: reductions over K correspondences on `st`, then out17 =
// {transformation (16), point-to-point rmse}.  corr == nullptr: rows are already gathered.
static int estimate_on_device(const double *d_src, const double *d_tgt, const double *d_nrm, const int *d_corr,
                              int64_t K, int estimator, const double *gravity_axis, cudaStream_t st, double out17[17]) {
    vb::SolveParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.estimator = estimator;
    sp.g[0] = 0.0; sp.g[1] = 1.0; sp.g[2] = 0.0;
    if (estimator == VB200_EST_P2PLANE_GRAVITY && gravity_axis) {
        double l = sqrt(gravity_axis[0] * gravity_axis[0] + gravity_axis[1] * gravity_axis[1] + gravity_axis[2] * gravity_axis[2]);
        if (!(l > 0.0)) return VB200_ERR_INVALID;
        for (int a = 0; a < 3; a++) sp.g[a] = gravity_axis[a] / l;
    }
    const int nblk = vb::div_up(K, vb::kChunk);
    vb::DevBuf<double> d_part(st), d_out(st);
    VB_CUDA(d_part.alloc((size_t)nblk * vb::kAcc));
    VB_CUDA(d_out.alloc(17));
    if (estimator != VB200_EST_P2P)
        vb::k_estimate<1><<<nblk, vb::kPassTpb, 0, st>>>(d_src, d_tgt, d_nrm, d_corr, K, d_part.p);
    else
        vb::k_estimate<0><<<nblk, vb::kPassTpb, 0, st>>>(d_src, d_tgt, d_nrm, d_corr, K, d_part.p);
    vb::k_estimate_solve<<<1, 256, 0, st>>>(d_part.p, nblk, d_src, d_corr, sp, d_out.p);
    VB_CUDA(cudaGetLastError());
    VB_CUDA(cudaMemcpyAsync(out17, d_out.p, sizeof(double) * 17, cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    return VB200_OK;
}

// host clouds: marshal the K referenced rows (the reference gathers them too, TransformationEstimation.cpp:52-57)
// and upload 72 bytes per correspondence instead of both clouds
static int estimate_from_host(const double *src_xyz, int64_t m, const double *tgt_xyz, const double *tgt_nrm, int64_t n,
                              const int32_t *corr, int64_t K, int estimator, const double *gravity_axis, int device,
                              double out17[17]) {
    const bool plane = estimator != VB200_EST_P2P;
    VB_TRY(vb::select_device(device));
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
    vb::DevBuf<double> d_in((cudaStream_t) nullptr);  // stream-ordered on the default stream (estimate_on_device runs there)
    VB_CUDA(d_in.alloc((size_t)K * 9));
    VB_CUDA(cudaMemcpy(d_in.p, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice));
    return estimate_on_device(d_in.p, d_in.p + 3 * K, d_in.p + 6 * K, nullptr, K, estimator, gravity_axis, nullptr, out17);
}

extern "C" int vb200_estimate(const double *src_xyz, int64_t m, const double *tgt_xyz, const double *tgt_nrm,
                              int64_t n, const int32_t *corr, int64_t K, int estimator, const double *gravity_axis,
                              int device, double out_T[16]) {
    if (!out_T || K < 0 || (K > 0 && (!src_xyz || !tgt_xyz || !corr))) return VB200_ERR_INVALID;
    if (estimator < VB200_EST_P2P || estimator > VB200_EST_P2PLANE_GRAVITY) return VB200_ERR_INVALID;
    for (int i = 0; i < 16; i++) out_T[i] = (i % 5 == 0) ? 1.0 : 0.0;
    // corres.empty() || !target.HasNormals() -> Identity (TransformationEstimation.cpp:51,79-80)
    if (K == 0 || (estimator != VB200_EST_P2P && !tgt_nrm)) return VB200_OK;
    double out17[17];
    VB_TRY(estimate_from_host(src_xyz, m, tgt_xyz, tgt_nrm, n, corr, K, estimator, gravity_axis, device, out17));
    memcpy(out_T, out17, sizeof(double) * 16);
    return VB200_OK;
}

extern "C" int vb200_estimate_device(const void *d_src_xyz, int64_t m, const void *d_tgt_xyz, const void *d_tgt_nrm,
                                     int64_t n, const void *d_corr, int64_t K, int estimator,
                                     const double *gravity_axis, int device, void *cuda_stream, double out_T[16]) {
    if (!out_T || K < 0 || m < 0 || n < 0 || (K > 0 && (!d_src_xyz || !d_tgt_xyz || !d_corr))) return VB200_ERR_INVALID;
    if (estimator < VB200_EST_P2P || estimator > VB200_EST_P2PLANE_GRAVITY) return VB200_ERR_INVALID;
    for (int i = 0; i < 16; i++) out_T[i] = (i % 5 == 0) ? 1.0 : 0.0;
    if (K == 0 || (estimator != VB200_EST_P2P && !d_tgt_nrm)) return VB200_OK;
    VB_TRY(vb::select_device(device));
    double out17[17];
    VB_TRY(estimate_on_device((const double *)d_src_xyz, (const double *)d_tgt_xyz, (const double *)d_tgt_nrm,
                              (const int *)d_corr, K, estimator, gravity_axis, (cudaStream_t)cuda_stream, out17));
    memcpy(out_T, out17, sizeof(double) * 16);
    return VB200_OK;
}

extern "C" int vb200_rmse(const double *src_xyz, int64_t m, const double *tgt_xyz, int64_t n, const int32_t *corr,
                          int64_t K, int device, double *out_rmse) {
    if (!out_rmse || K < 0 || (K > 0 && (!src_xyz || !tgt_xyz || !corr))) return VB200_ERR_INVALID;
    *out_rmse = 0.0;  // corres.empty() -> 0 (src/constrained_ICP.cpp:17)
    if (K == 0) return VB200_OK;
    double out17[17];
    VB_TRY(estimate_from_host(src_xyz, m, tgt_xyz, nullptr, n, corr, K, VB200_EST_P2P, nullptr, device, out17));
    *out_rmse = out17[16];
    return VB200_OK;
}

extern "C" int vb200_batch_set_option(vb200_batch_t *batch, int option, int value) {
    if (!batch) return VB200_ERR_INVALID;
    Batch *b = reinterpret_cast<Batch *>(batch);
    switch (option) {
        case VB200_OPT_NN_CACHE: b->use_cache = value != 0; return VB200_OK;
        case VB200_OPT_SPLIT_TIMING: b->split_timing = value != 0; return VB200_OK;
        default: return VB200_ERR_INVALID;
    }
}

#ifdef VB_STATS
// dev builds only (scripts/build_variants.sh ... "-DVB_STATS"): search-path counters of grid.cuh
extern "C" int vb200_debug_stats(unsigned long long *out, int reset) {
    if (out) cudaMemcpyFromSymbol(out, vb::g_stats, sizeof(unsigned long long) * 24);
    if (reset) {
        unsigned long long z[24] = {0};
        cudaMemcpyToSymbol(vb::g_stats, z, sizeof(z));
    }
    return 0;
}
extern "C" int vb200_debug_states(vb200_batch_t *batch, float *cum, float *last_move) {
    Batch *b = reinterpret_cast<Batch *>(batch);
    std::vector<vb::ProbState> st((size_t)b->P);
    cudaMemcpy(st.data(), b->d_states, sizeof(vb::ProbState) * (size_t)b->P, cudaMemcpyDeviceToHost);
    for (int p = 0; p < b->P; p++) { cum[p] = st[p].cum; last_move[p] = st[p].last_move; }
    return 0;
}
#endif
