// knn.cu — vb200_knn1 / vb200_knn1_device: radius-bounded 1-NN for a batch of query points.
// Replaces the per-point KDTreeFlann::SearchHybrid(query, radius, 1) loop of
// GetRegistrationResultAndCorrespondences (O3D/src/Core/Registration/Registration.cpp:62-72,
// O3D/src/Core/Geometry/KDTreeFlann.cpp:165-189).  Results are bit-identical to the reference's double
// arithmetic: same neighbour index, same d2 (see grid.cuh).
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <cmath>

#include "scene.cuh"

namespace vb {

namespace {

constexpr int kTpb = 256;

__global__ void __launch_bounds__(kTpb) k_knn1(GridDev G, const double *__restrict__ q, int64_t nq,
                                               const int *__restrict__ perm, double r2, float r2_ub,
                                               int *__restrict__ out_idx, double *__restrict__ out_d2) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = s < nq;  // whole warps stay alive: the search is warp-cooperative
    const int64_t i = live ? perm[s] : 0;  // queries are visited in grid order, results land in caller order
    double x = 0.0, y = 0.0, z = 0.0, d2 = 0.0;
    QueryCtx c;
    bool inside = false;
    if (live) {
        x = q[3 * i]; y = q[3 * i + 1]; z = q[3 * i + 2];
        inside = make_query(G.p, x, y, z, c);
    }
    __shared__ LaneRuns<kTpb> runs;
    const int bs = nn_search_hybrid<kTpb>(G, inside, c, x, y, z, r2, r2_ub, -1, runs, &d2);
    if (live) {
        out_idx[i] = bs >= 0 ? __ldg(G.orig + bs) : -1;
        out_d2[i] = bs >= 0 ? d2 : 0.0;
    }
}


// one warp per query (few / scattered queries)
__global__ void __launch_bounds__(kTpb) k_knn1_wpq(GridDev G, const double *__restrict__ q, int64_t nq, double r2,
                                                   float r2_ub, int bfs, int *__restrict__ out_idx,
                                                   double *__restrict__ out_d2) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= nq) return;  // warp-uniform exit
    const double x = q[3 * i], y = q[3 * i + 1], z = q[3 * i + 2];
    QueryCtx c;
    int bs = -1;
    double d2 = 0.0;
    __shared__ WpqScratch ws[kTpb / 32];
    if (make_query(G.p, x, y, z, c)) {
        bs = bfs ? nn_search_wpq_bfs(G, c, x, y, z, r2, r2_ub, ws[threadIdx.x >> 5], &d2)
                 : nn_search_wpq(G, c, x, y, z, r2, r2_ub, &d2);
    }
    if ((threadIdx.x & 31) == 0) {
        out_idx[i] = bs >= 0 ? __ldg(G.orig + bs) : -1;
        out_d2[i] = bs >= 0 ? d2 : 0.0;
    }
}

// ---- exhaustive search: vb200_knn1_bruteforce -------------------------------------------------------
// The same operator without any index: every (query, target) distance in the reference's double arithmetic.
// It needs no scene handle (nothing to build), streams the target cloud exactly once per chunk of 8 queries,
// and is the independent on-device check of the grid search at sizes no CPU oracle finishes
// (tests/test_gpu_knn.py).  With a handful of queries it is the HBM-streaming case SURVEY §8d's algorithmic
// bytes describe (24 B per target point here: the cloud stays in the caller's f64 layout); beyond ~8 queries
// it is bound by the FP64 pipe (8 rounded operations per pair, no FMA: FLANN's operation order).
//
// grid = (slices of the target cloud, chunks of up to 8 queries).  A block streams its slice through shared
// memory in tiles of kBfTile points moved by the TMA engine (cp.async.bulk + mbarrier, kBfStages tiles in
// flight, issued by one thread); thread i takes points i, i + 128, ... of a tile against the chunk's queries
// held in registers; at the end the block's per-query (d2, index) minima are combined with warp shuffles and
// one shared-memory step, and k_bf_merge combines the slices.  Ties: lowest target index, as everywhere.
constexpr int kBfTpb = 128, kBfTile = 1024, kBfStages = 3;
constexpr int kBfTileBytes = kBfTile * 24;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
// one elected thread: arm the barrier with the byte count, then let the TMA engine copy global -> shared
__device__ __forceinline__ void tma_load_1d(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// (d, i) before (bd, bi): smaller distance, then lower index.  The distances are sums of squares (>= +0.0, or
// NaN), for which the IEEE bit patterns order like signed integers — the comparison runs on the integer pipe
// and leaves the FP64 pipe to the 8 operations of the distance itself.  NaN patterns sort above everything
// (a NaN query or target point is never anyone's neighbour, as in the reference: `dist < worst_dist` is false).
__device__ __forceinline__ bool bf_less(double d, int i, double bd, int bi) {
    long long a = __double_as_longlong(d);
    const long long b = __double_as_longlong(bd);
    if (a < 0) a = 0x7fffffffffffffffll;  // a NaN with its sign bit set (a sum of squares is never negative)
    return a < b || (a == b && i < bi);
}

template <int NQ>
__global__ void __launch_bounds__(kBfTpb) k_bf_knn1(const double *__restrict__ tgt, int64_t n, int64_t slice_pts,
                                                    const double *__restrict__ q, int64_t nq,
                                                    double *__restrict__ part_d2, int *__restrict__ part_idx) {
    extern __shared__ __align__(128) unsigned char bf_smem[];
    double *tiles = reinterpret_cast<double *>(bf_smem);  // kBfStages x kBfTile x 3 doubles
    __shared__ __align__(8) unsigned long long bars[kBfStages];
    __shared__ double red_d[kBfTpb / 32][NQ];
    __shared__ int red_i[kBfTpb / 32][NQ];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t p_begin = (int64_t)blockIdx.x * slice_pts;
    const int64_t p_end = min(n, p_begin + slice_pts);
    const int64_t q0 = (int64_t)blockIdx.y * NQ;
    // the chunk's queries: warp-uniform loads, kept in registers (a missing query repeats the last one)
    double qx[NQ], qy[NQ], qz[NQ], bd[NQ];
    int bi[NQ];
#pragma unroll
    for (int j = 0; j < NQ; j++) {
        const int64_t qi = min(q0 + j, nq - 1);
        qx[j] = q[3 * qi]; qy[j] = q[3 * qi + 1]; qz[j] = q[3 * qi + 2];
        bd[j] = 1.0e300; bi[j] = 0x7fffffff;
    }
    const int64_t npts = p_end - p_begin;
    const int nfull = (int)(npts / kBfTile);  // whole tiles go through the TMA engine, the ragged tail through plain loads
    if (tid == 0) {
        for (int s2 = 0; s2 < kBfStages; s2++) mbar_init(smem_u32(&bars[s2]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int t = 0; t < kBfStages && t < nfull; t++)
            tma_load_1d(smem_u32(tiles + (size_t)t * kBfTile * 3), tgt + 3 * (p_begin + (int64_t)t * kBfTile), kBfTileBytes,
                        smem_u32(&bars[t]));
    }
    for (int t = 0; t < nfull; t++) {
        const int st = t % kBfStages;
        mbar_wait(smem_u32(&bars[st]), (unsigned)((t / kBfStages) & 1));
        const double *tile = tiles + (size_t)st * kBfTile * 3;
        const int base = (int)(p_begin + (int64_t)t * kBfTile);
#pragma unroll 2
        for (int i = tid; i < kBfTile; i += kBfTpb) {
            const double tx = tile[3 * i], ty = tile[3 * i + 1], tz = tile[3 * i + 2];
#pragma unroll
            for (int j = 0; j < NQ; j++) {
                const double d = l2_exact(qx[j], qy[j], qz[j], tx, ty, tz);
                if (bf_less(d, base + i, bd[j], bi[j])) { bd[j] = d; bi[j] = base + i; }
            }
        }
        __syncthreads();  // every thread is done with this stage: it can be refilled
        if (tid == 0 && t + kBfStages < nfull)
            tma_load_1d(smem_u32(tiles + (size_t)st * kBfTile * 3),
                        tgt + 3 * (p_begin + (int64_t)(t + kBfStages) * kBfTile), kBfTileBytes, smem_u32(&bars[st]));
    }
    for (int64_t i = p_begin + (int64_t)nfull * kBfTile + tid; i < p_end; i += kBfTpb) {
        const double tx = tgt[3 * i], ty = tgt[3 * i + 1], tz = tgt[3 * i + 2];
#pragma unroll
        for (int j = 0; j < NQ; j++) {
            const double d = l2_exact(qx[j], qy[j], qz[j], tx, ty, tz);
            if (bf_less(d, (int)i, bd[j], bi[j])) { bd[j] = d; bi[j] = (int)i; }
        }
    }
    // per-query minimum over the block: butterflies inside each warp, then one step through shared memory
#pragma unroll
    for (int j = 0; j < NQ; j++) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bd[j], o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi[j], o);
            if (bf_less(od, oi, bd[j], bi[j])) { bd[j] = od; bi[j] = oi; }
        }
        if (lane == 0) { red_d[warp][j] = bd[j]; red_i[warp][j] = bi[j]; }
    }
    __syncthreads();
    if (tid < NQ && q0 + tid < nq) {
        double d = red_d[0][tid];
        int i = red_i[0][tid];
        for (int w = 1; w < kBfTpb / 32; w++)
            if (bf_less(red_d[w][tid], red_i[w][tid], d, i)) { d = red_d[w][tid]; i = red_i[w][tid]; }
        part_d2[(int64_t)blockIdx.x * nq + q0 + tid] = d;
        part_idx[(int64_t)blockIdx.x * nq + q0 + tid] = i;
    }
}

// one warp per query: the lanes stride over the slices' minima (one thread walking 296 dependent loads took
// longer than the search itself), then a butterfly
__global__ void __launch_bounds__(256) k_bf_merge(const double *__restrict__ part_d2, const int *__restrict__ part_idx,
                                                  int nslices, int64_t nq, double r2, int *__restrict__ out_idx,
                                                  double *__restrict__ out_d2) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= nq) return;  // warp-uniform
    double d = 1.0e300;
    int b = 0x7fffffff;
    for (int s2 = lane; s2 < nslices; s2 += 32) {
        const double od = part_d2[(int64_t)s2 * nq + i];
        const int oi = part_idx[(int64_t)s2 * nq + i];
        if (bf_less(od, oi, d, b)) { d = od; b = oi; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, d, o);
        const int oi = __shfl_xor_sync(0xffffffffu, b, o);
        if (bf_less(od, oi, d, b)) { d = od; b = oi; }
    }
    if (lane == 0) {
        const bool hit = b != 0x7fffffff && d < r2;  // accepted iff d2 < (double)(float)(r*r) (KDTreeFlann.cpp:185)
        out_idx[i] = hit ? b : -1;
        out_d2[i] = hit ? d : 0.0;
    }
}

// ---- exhaustive search, many queries: f32 screening of a packed copy ----------------------------------
// From a dozen queries up the f64 kernel above is bound by the FP64 pipe (8 rounded operations per pair).  This
// one screens every pair in f32 and takes in double only the pairs that can matter, so the answer is the same
// bit for bit:
//   * k_bf_pack writes the cloud once as float4 tiles: the point centred on the middle of the bounding box of
//     cloud and queries, rounded to f32, and w = |t~|^2 (16 B per point, padded with NaN to whole tiles);
//   * a block holds 8 x kBfQChunks queries (8 per thread, as -2 q~ in registers) and streams its slice of the
//     packed cloud through shared memory with the TMA engine (the same mbarrier ring as above); a pair costs
//     three FFMA and a compare: s = w - 2 q~.t~ (= |q~ - t~|^2 - |q~|^2 up to rounding) against a per-query
//     threshold c;
//   * c is rigorous: with e the per-axis error of q~ - t~ against q - t and E the rounding of the three FFMA,
//     every pair with true d2 <= B has s < (sqrt(B) + sqrt(3) e)^2 (1 + 1e-6) + E - |q~|^2 =: c(B).  B starts at
//     the acceptance threshold r2 and drops to the thread's exact best; a pair below c is evaluated exactly
//     (FLANN's operation order, from the caller's f64 arrays) and competes on (d2, index) as above.
//   NaN points or queries give NaN sums: never below c, never anyone's neighbour.  A cloud too large or too far
//   from finite for the bound to be useful (E of the order of r2) takes the f64 kernel instead.
constexpr int kBfPackTile = 1024;                    // points per packed tile (16 KB)
constexpr int kBfPackStages = 3;
constexpr int kBfTgtLanes = 64;                      // threads striding over a tile's points
constexpr int kBfQChunks = 4;                        // x 8 queries each = 32 queries per block
constexpr int kBfF32Tpb = kBfTgtLanes * kBfQChunks;  // 256 screening threads; the block has one more warp, the tile producer
constexpr int kBfF32Block = kBfF32Tpb + 32;
constexpr int kBfNQ = 8;
constexpr int kBfQPerBlock = kBfNQ * kBfQChunks;

struct BfScreen {
    double ctr[3];   // subtracted before rounding to f32
    double sqrt3e;   // sqrt(3) x per-axis error bound of (q~ - t~) against (q - t)
    double E;        // rounding bound of the computed sum
    double r2;       // (double)(float)(radius^2)
};

__global__ void __launch_bounds__(256) k_bf_pack(const double *__restrict__ tgt, int64_t n, int64_t n_pad, BfScreen sc,
                                                 float4 *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    const float nan = __int_as_float(0x7fc00000);
    float4 v = make_float4(nan, nan, nan, nan);
    if (i < n) {
        v.x = (float)(tgt[3 * i] - sc.ctr[0]);
        v.y = (float)(tgt[3 * i + 1] - sc.ctr[1]);
        v.z = (float)(tgt[3 * i + 2] - sc.ctr[2]);
        v.w = (float)((double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z);
    }
    out[i] = v;
}

// c(B): see above.  Rounded up, then one more ulp-sized step for the conversion itself.
__device__ __forceinline__ float bf_threshold(const BfScreen &sc, double B, double qq) {
    const double r = sqrt(B) + sc.sqrt3e;
    const float c = __double2float_ru(r * r * (1.0 + 1e-6) + sc.E - qq);
    return c + fabsf(c) * 1.2e-7f + 1e-37f;
}

// The exact evaluation of the pairs of one target point that passed the screen (rare: kept out of line, the running
// bests in shared memory, so that the screening loop keeps its registers and its single branch per point).  Returns
// the chunk's thresholds, lowered where a pair became its query's best.
struct BfF8 {
    float v[kBfNQ];
};
// float <-> int with the order preserved (thresholds can be negative: they live in s-space, c = bound - |q~|^2), so
// that the lowest threshold of a query is kept with an integer atomicMin — in shared memory for the block, in global
// memory for every block scanning another slice of the cloud for the same query
__device__ __forceinline__ int bf_key(float c) { const int b = __float_as_int(c); return b >= 0 ? b : b ^ 0x7fffffff; }
__device__ __forceinline__ float bf_unkey(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }
static __device__ __noinline__ BfF8 bf_exact_point(const BfScreen *sc, const double *__restrict__ q, const double *__restrict__ tgt,
                                                   int64_t q0, int64_t nq, int idx, double *bd, int *bi, int *c_blk, int *c_all,
                                                   BfF8 sum, BfF8 c, BfF8 qx2, BfF8 qy2, BfF8 qz2) {
    const double tx = tgt[3 * (int64_t)idx], ty = tgt[3 * (int64_t)idx + 1], tz = tgt[3 * (int64_t)idx + 2];
#pragma unroll 1
    for (int j = 0; j < kBfNQ; j++) {
        if (!(sum.v[j] < c.v[j])) continue;
        const int64_t qi = min(q0 + j, nq - 1);
        const double d = l2_exact(q[3 * qi], q[3 * qi + 1], q[3 * qi + 2], tx, ty, tz);
        double *bdj = bd + (size_t)j * kBfF32Tpb;  // [query][thread] columns
        int *bij = bi + (size_t)j * kBfF32Tpb;
        if (!bf_less(d, idx, *bdj, *bij)) continue;
        *bdj = d;
        *bij = idx;
        const float x = -0.5f * qx2.v[j], y = -0.5f * qy2.v[j], z = -0.5f * qz2.v[j];
        const float cn = bf_threshold(*sc, d, (double)x * x + (double)y * y + (double)z * z);
        if (cn < c.v[j]) {
            // every thread and every block working on this query may use the bound: a real point lies this close
            c.v[j] = cn;
            atomicMin(c_blk + j, bf_key(cn));
            if (q0 + j < nq) atomicMin(c_all + q0 + j, bf_key(cn));
        }
    }
    return c;
}

__global__ void __launch_bounds__(kBfF32Block, 3) k_bf_knn1_f32(const float4 *__restrict__ pack, int64_t slice_pts,
                                                             int64_t n_pad, const double *__restrict__ tgt,
                                                             const double *__restrict__ q, int64_t nq, BfScreen sc,
                                                             int *__restrict__ c_all, double *__restrict__ part_d2,
                                                             int *__restrict__ part_idx) {
    extern __shared__ __align__(128) unsigned char bf_smem[];
    float4 *tiles = reinterpret_cast<float4 *>(bf_smem);  // kBfPackStages x kBfPackTile
    // The lowest threshold known for each of the block's queries (order-preserving keys).  Without sharing, every
    // point inside the radius (~1e-4 of the cloud) would cost an exact evaluation — two dependent global loads — on
    // whichever thread met it, and the block would wait for that thread at every tile; with it a query's threshold
    // drops to its true neighbourhood as soon as ANY thread of ANY block has seen a close point.
    __shared__ int c_blk[kBfQPerBlock];
    // full[s]: the TMA engine has landed stage s;  empty[s]: every screening warp is done with it.  The warps never
    // wait for each other (a block-wide barrier per tile left them idle a third of the time: whichever warp met an
    // exact evaluation held up the rest); the producer warp refills a stage as soon as its last reader has left.
    __shared__ __align__(8) unsigned long long bars[kBfPackStages], empty[kBfPackStages];
    __shared__ double red_d[kBfF32Tpb / 32][kBfNQ];
    __shared__ int red_i[kBfF32Tpb / 32][kBfNQ];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tl = tid & (kBfTgtLanes - 1), qc = tid / kBfTgtLanes;
    const int64_t p_begin = (int64_t)blockIdx.x * slice_pts;
    const int64_t p_end = min(n_pad, p_begin + slice_pts);
    const int64_t q0 = (int64_t)blockIdx.y * kBfQPerBlock + (int64_t)qc * kBfNQ;
    BfF8 qx2, qy2, qz2, c;
    // the running exact best per (thread, query): touched only by the rare exact evaluations
    __shared__ double bd_sh[kBfNQ][kBfF32Tpb];
    __shared__ int bi_sh[kBfNQ][kBfF32Tpb];
    __shared__ BfScreen sc_sh;
    if (tid == 0) sc_sh = sc;
    const bool producer = tid >= kBfF32Tpb;  // (warp-uniform)
    if (!producer) {
#pragma unroll
        for (int j = 0; j < kBfNQ; j++) {
            const int64_t qi = min(q0 + j, nq - 1);  // (a missing query repeats the last one; its result is not written)
            const float x = (float)(q[3 * qi] - sc.ctr[0]), y = (float)(q[3 * qi + 1] - sc.ctr[1]),
                        z = (float)(q[3 * qi + 2] - sc.ctr[2]);
            qx2.v[j] = -2.0f * x; qy2.v[j] = -2.0f * y; qz2.v[j] = -2.0f * z;
            c.v[j] = bf_threshold(sc, sc.r2, (double)x * x + (double)y * y + (double)z * z);
            bd_sh[j][tid] = 1.0e300; bi_sh[j][tid] = 0x7fffffff;
            if (tl == 0) c_blk[qc * kBfNQ + j] = bf_key(c.v[j]);
        }
    }
    const int ntiles = (int)((p_end - p_begin) / kBfPackTile);  // slices are whole tiles
    if (tid == 0) {
        for (int s2 = 0; s2 < kBfPackStages; s2++) {
            mbar_init(smem_u32(&bars[s2]), 1);
            mbar_init(smem_u32(&empty[s2]), kBfF32Tpb / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    constexpr unsigned kTileBytes = kBfPackTile * sizeof(float4);
    if (producer) {
        // the producer warp: lane 0 keeps the ring full, the others bring in the thresholds other blocks have found for
        // this block's queries (stale values are merely higher: safe)
        for (int t = 0; t < ntiles; t++) {
            const int st = t % kBfPackStages;
            if (lane == 0) {
                if (t >= kBfPackStages) mbar_wait(smem_u32(&empty[st]), (unsigned)(((t / kBfPackStages) - 1) & 1));
                tma_load_1d(smem_u32(tiles + (size_t)st * kBfPackTile), pack + p_begin + (int64_t)t * kBfPackTile, kTileBytes,
                            smem_u32(&bars[st]));
            }
            const int64_t qg = (int64_t)blockIdx.y * kBfQPerBlock + lane;
            if (lane < kBfQPerBlock && qg < nq) atomicMin(&c_blk[lane], __ldcg(c_all + qg));
            __syncwarp();
        }
    } else {
        for (int t = 0; t < ntiles; t++) {
            const int st = t % kBfPackStages;
            mbar_wait(smem_u32(&bars[st]), (unsigned)((t / kBfPackStages) & 1));
            const float4 *tile = tiles + (size_t)st * kBfPackTile;
            const int base = (int)(p_begin + (int64_t)t * kBfPackTile);
#pragma unroll
            for (int j = 0; j < kBfNQ; j++) c.v[j] = fminf(c.v[j], bf_unkey(c_blk[qc * kBfNQ + j]));
            // the screen: per point eight independent chains of three FFMA and eight compares folded into one
            // predicate; a point with a pair below its threshold (rare) only sets its bit — the exact evaluations, and
            // the call they need, stay out of this loop
            static_assert(kBfPackTile / kBfTgtLanes <= 32, "one bit per point of the thread's share of a tile");
            unsigned hit = 0u, bit = 1u;
#pragma unroll 4
            for (int k = 0; k < kBfPackTile / kBfTgtLanes; k++, bit <<= 1) {
                const float4 p = tile[tl + k * kBfTgtLanes];
                float sum[kBfNQ];
#pragma unroll
                for (int j = 0; j < kBfNQ; j++) sum[j] = fmaf(qz2.v[j], p.z, fmaf(qy2.v[j], p.y, fmaf(qx2.v[j], p.x, p.w)));
                // hit |= bit if any sum is below its threshold: eight compares OR-ed into ONE predicate (setp.lt.or), one
                // predicated OR.  (Written in C the compiler materialises every compare with a select: 5 instructions a
                // pair instead of 4.)
                static_assert(kBfNQ == 8, "the predicate chain below is written out for 8 queries");
                asm("{\n"
                    ".reg .pred p;\n"
                    "setp.lt.f32 p, %2, %10;\n"
                    "setp.lt.or.f32 p, %3, %11, p;\n"
                    "setp.lt.or.f32 p, %4, %12, p;\n"
                    "setp.lt.or.f32 p, %5, %13, p;\n"
                    "setp.lt.or.f32 p, %6, %14, p;\n"
                    "setp.lt.or.f32 p, %7, %15, p;\n"
                    "setp.lt.or.f32 p, %8, %16, p;\n"
                    "setp.lt.or.f32 p, %9, %17, p;\n"
                    "@p or.b32 %0, %0, %1;\n"
                    "}"
                    : "+r"(hit)
                    : "r"(bit), "f"(sum[0]), "f"(sum[1]), "f"(sum[2]), "f"(sum[3]), "f"(sum[4]), "f"(sum[5]), "f"(sum[6]), "f"(sum[7]),
                      "f"(c.v[0]), "f"(c.v[1]), "f"(c.v[2]), "f"(c.v[3]), "f"(c.v[4]), "f"(c.v[5]), "f"(c.v[6]), "f"(c.v[7]));
            }
            while (hit) {
                const int k = __ffs(hit) - 1;
                hit &= hit - 1u;
                const float4 p = tile[tl + k * kBfTgtLanes];
                BfF8 sum;
#pragma unroll
                for (int j = 0; j < kBfNQ; j++) sum.v[j] = fmaf(qz2.v[j], p.z, fmaf(qy2.v[j], p.y, fmaf(qx2.v[j], p.x, p.w)));
                c = bf_exact_point(&sc_sh, q, tgt, q0, nq, base + tl + k * kBfTgtLanes, &bd_sh[0][tid], &bi_sh[0][tid],
                                   &c_blk[qc * kBfNQ], c_all, sum, c, qx2, qy2, qz2);
            }
            __syncwarp();  // the warp is done with this stage
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[st])) : "memory");
        }
    }
    __syncthreads();
    // per-query minimum over the block's threads that share the query chunk: butterflies inside each warp, then the
    // chunk's two warps through shared memory
#pragma unroll
    for (int j = 0; j < kBfNQ && !producer; j++) {
        double d = bd_sh[j][tid];
        int b = bi_sh[j][tid];
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, d, o);
            const int oi = __shfl_xor_sync(0xffffffffu, b, o);
            if (bf_less(od, oi, d, b)) { d = od; b = oi; }
        }
        if (lane == 0) { red_d[warp][j] = d; red_i[warp][j] = b; }
    }
    __syncthreads();
    if (!producer && tl < kBfNQ && q0 + tl < nq) {
        constexpr int wpc = kBfTgtLanes / 32;  // warps per query chunk
        double d = red_d[qc * wpc][tl];
        int i = red_i[qc * wpc][tl];
        for (int w = 1; w < wpc; w++)
            if (bf_less(red_d[qc * wpc + w][tl], red_i[qc * wpc + w][tl], d, i)) { d = red_d[qc * wpc + w][tl]; i = red_i[qc * wpc + w][tl]; }
        part_d2[(int64_t)blockIdx.x * nq + q0 + tl] = d;
        part_idx[(int64_t)blockIdx.x * nq + q0 + tl] = i;
    }
}

// the screened path; *done = false when the inputs do not allow it (the caller then runs the f64 kernel)
int bf_launch_f32(const double *d_tgt, int64_t n, const double *d_q, int64_t nq, double r2, int *d_idx, double *d_d2,
                  cudaStream_t st, bool *done) {
    *done = false;
    double lo[3], hi[3], qlo[3], qhi[3];
    VB_TRY(device_bbox(d_tgt, n, lo, hi, st));  // (fmin / fmax: NaN coordinates are ignored)
    VB_TRY(device_bbox(d_q, nq, qlo, qhi, st));
    BfScreen sc;
    double cmax = 0.0, cabs = 0.0;
    for (int a = 0; a < 3; a++) {
        const double l = std::min(lo[a], qlo[a]), h = std::max(hi[a], qhi[a]);
        if (!(l <= h) || !std::isfinite(l) || !std::isfinite(h)) return VB200_OK;  // all-NaN or infinite input
        sc.ctr[a] = 0.5 * (l + h);
        cmax = std::max(cmax, 0.5 * (h - l));
        cabs = std::max(cabs, fabs(sc.ctr[a]));
    }
    const double u = 5.9604644775390625e-8;  // 2^-24
    // per axis: |t~ - (t - c)| <= u |t - c| + rounding of the f64 subtraction, the same for q~; then the f32 sum:
    // w = |t~|^2 rounded once and three FFMA whose results stay below |t~|^2 + 2 |q~| |t~| <= 9 cmax^2 (1 + ...)
    const double e = 2.02 * u * cmax + 4e-16 * (cmax + cabs);
    sc.sqrt3e = 1.7320508075688772 * e * 1.001;
    sc.E = 64.0 * u * cmax * cmax + 1e-300;
    sc.r2 = r2;
    if (!(sc.E < 0.25 * r2) || !(cmax < 1e15)) return VB200_OK;  // the screen would pass too much: f64 kernel
    const int64_t n_pad = (int64_t)div_up(n, kBfPackTile) * kBfPackTile;
    DevBuf<float4> d_pack(st);
    VB_CUDA(d_pack.alloc((size_t)n_pad));
    k_bf_pack<<<div_up(n_pad, 256), 256, 0, st>>>(d_tgt, n, n_pad, sc, d_pack.p);
    VB_CUDA(cudaGetLastError());
    const int nqb = div_up(nq, kBfQPerBlock);
    if (nqb > 65535) return VB200_OK;
    // slices of whole tiles; enough blocks for ~8 waves of 2 blocks per SM so that the last wave costs little
    const int64_t ntile = n_pad / kBfPackTile;
    int64_t nslices = std::max<int64_t>(1, std::min<int64_t>(div_up(ntile, 2), div_up(16 * kNumSMsB200, nqb)));
    const int64_t slice_pts = (int64_t)div_up(ntile, nslices) * kBfPackTile;
    nslices = div_up(n_pad, slice_pts);
    DevBuf<double> p_d2(st);
    DevBuf<int> p_idx(st);
    VB_CUDA(p_d2.alloc((size_t)nslices * (size_t)nq));
    VB_CUDA(p_idx.alloc((size_t)nslices * (size_t)nq));
    DevBuf<int> d_call(st);
    VB_CUDA(d_call.alloc((size_t)nq));
    VB_CUDA(cudaMemsetAsync(d_call.p, 0x7f, sizeof(int) * (size_t)nq, st));  // keys of 3.4e38: above any threshold
    const size_t smem = (size_t)kBfPackStages * kBfPackTile * sizeof(float4);
    static bool opted[64] = {};
    int dev = 0;
    VB_CUDA(cudaGetDevice(&dev));
    if (dev >= 64 || !opted[dev]) {
        VB_CUDA(cudaFuncSetAttribute(k_bf_knn1_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev < 64) opted[dev] = true;
    }
    k_bf_knn1_f32<<<dim3((unsigned)nslices, (unsigned)nqb), kBfF32Block, smem, st>>>(d_pack.p, slice_pts, n_pad, d_tgt, d_q, nq, sc,
                                                                                 d_call.p, p_d2.p, p_idx.p);
    k_bf_merge<<<div_up(nq * 32, 256), 256, 0, st>>>(p_d2.p, p_idx.p, (int)nslices, nq, r2, d_idx, d_d2);
    VB_CUDA(cudaGetLastError());
    *done = true;
    return VB200_OK;
}

int bf_launch(const double *d_tgt, int64_t n, const double *d_q, int64_t nq, double radius, int *d_idx, double *d_d2,
              cudaStream_t st) {
    if (!(radius > 0.0) || n < 0 || nq < 0 || n > 0x7ffffffe || nq > 0x7fffffff) return VB200_ERR_INVALID;
    if ((reinterpret_cast<uintptr_t>(d_tgt) & 15) != 0) return VB200_ERR_INVALID;  // the bulk copies need 16-byte alignment
    if (nq == 0) return VB200_OK;
    const double r2 = (double)(float)(radius * radius);
    // a dozen queries or more: the f32-screened kernel (the f64 one below is FP64-bound from ~8 queries)
    if (nq >= 16 && n >= 1) {
        bool done = false;
        VB_TRY(bf_launch_f32(d_tgt, n, d_q, nq, r2, d_idx, d_d2, st, &done));
        if (done) return VB200_OK;
    }
    // queries per block: 8, or fewer when there are fewer (1-2 queries leave the kernel bound by the HBM stream,
    // 8 by the FP64 pipe)
    const int cq = nq >= 5 ? 8 : nq >= 3 ? 4 : (int)nq;
    const int nchunks = div_up(nq, cq);
    if (nchunks > 65535) return VB200_ERR_INVALID;  // grid.y; 524 280 queries per call
    // enough blocks for two waves when there are few queries; slices are whole tiles so every bulk copy starts
    // on a 16-byte boundary (1024 points x 24 B)
    int64_t nslices = std::max<int64_t>(1, std::min<int64_t>(div_up(2 * kNumSMsB200, nchunks), div_up(std::max<int64_t>(n, 1), 4 * kBfTile)));
    int64_t slice_pts = div_up(div_up(std::max<int64_t>(n, 1), nslices), kBfTile) * (int64_t)kBfTile;
    nslices = std::max<int64_t>(1, div_up(std::max<int64_t>(n, 1), slice_pts));
    DevBuf<double> p_d2(st);
    DevBuf<int> p_idx(st);
    VB_CUDA(p_d2.alloc((size_t)nslices * (size_t)nq));
    VB_CUDA(p_idx.alloc((size_t)nslices * (size_t)nq));
    const size_t smem = (size_t)kBfStages * kBfTileBytes;
    // the opt-in for > 48 KB of dynamic shared memory is per device and per kernel: once each
    static bool opted[64][4] = {};
    int dev = 0;
    VB_CUDA(cudaGetDevice(&dev));
    const int variant = cq == 1 ? 0 : cq == 2 ? 1 : cq == 4 ? 2 : 3;
    auto launch = [&](auto kernel) -> cudaError_t {
        if (dev >= 64 || !opted[dev][variant]) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            if (dev < 64) opted[dev][variant] = true;
        }
        kernel<<<dim3((unsigned)nslices, (unsigned)nchunks), kBfTpb, smem, st>>>(d_tgt, n, slice_pts, d_q, nq, p_d2.p, p_idx.p);
        return cudaSuccess;
    };
    switch (cq) {
        case 1: VB_CUDA(launch(k_bf_knn1<1>)); break;
        case 2: VB_CUDA(launch(k_bf_knn1<2>)); break;
        case 4: VB_CUDA(launch(k_bf_knn1<4>)); break;
        default: VB_CUDA(launch(k_bf_knn1<8>)); break;
    }
    k_bf_merge<<<div_up(nq * 32, 256), 256, 0, st>>>(p_d2.p, p_idx.p, (int)nslices, nq, r2, d_idx, d_d2);
    VB_CUDA(cudaGetLastError());
    return VB200_OK;
}

constexpr int64_t kWarpPerQueryMax = 262144;  // below this many queries a warp per query fills the GPU better

int knn1_launch(Scene *sc, const double *d_q, int64_t nq, double radius, int *d_idx, double *d_d2) {
    if (!(radius > 0.0) || radius > sc->grid.p.cell * (1.0 + 1e-12)) return VB200_ERR_INVALID;
    if (nq == 0) return VB200_OK;
    if (nq > 0x7fffffff) return VB200_ERR_INVALID;
    const double r2 = (double)(float)(radius * radius);  // KDTreeFlann.cpp:185
    if (nq <= kWarpPerQueryMax) {
        k_knn1_wpq<<<div_up(nq * 32, kTpb), kTpb, 0, sc->stream>>>(sc->grid, d_q, nq, r2,
                                                                   r2_upper_bound(sc->grid.p, r2), 1, d_idx, d_d2);
        VB_CUDA(cudaGetLastError());
        return VB200_OK;
    }
    DevBuf<int> d_perm(sc->stream);
    VB_CUDA(d_perm.alloc((size_t)nq));
    VB_TRY(grid_order_points(sc, d_q, nq, d_perm.p));
    k_knn1<<<div_up(nq, kTpb), kTpb, 0, sc->stream>>>(sc->grid, d_q, nq, d_perm.p, r2,
                                                      r2_upper_bound(sc->grid.p, r2), d_idx, d_d2);
    VB_CUDA(cudaGetLastError());
    return VB200_OK;
}

}  // namespace

}  // namespace vb

using vb::Scene;

extern "C" int vb200_knn1_device(vb200_scene_t *scene, const void *d_q_xyz, int64_t Q, double radius,
                                 void *d_out_idx, void *d_out_d2) {
    if (!scene || Q < 0 || (Q > 0 && (!d_q_xyz || !d_out_idx || !d_out_d2))) return VB200_ERR_INVALID;
    Scene *sc = reinterpret_cast<Scene *>(scene);
    VB_CUDA(cudaSetDevice(sc->device));
    return vb::knn1_launch(sc, (const double *)d_q_xyz, Q, radius, (int *)d_out_idx, (double *)d_out_d2);
}

extern "C" int vb200_knn1(vb200_scene_t *scene, const double *q_xyz, int64_t Q, double radius, int32_t *out_idx,
                          double *out_d2) {
    if (!scene || Q < 0 || (Q > 0 && (!q_xyz || !out_idx || !out_d2))) return VB200_ERR_INVALID;
    Scene *sc = reinterpret_cast<Scene *>(scene);
    VB_CUDA(cudaSetDevice(sc->device));
    if (!(radius > 0.0) || radius > sc->grid.p.cell * (1.0 + 1e-12)) return VB200_ERR_INVALID;
    if (Q == 0) return VB200_OK;
    vb::DevBuf<double> d_q(sc->stream), d_d2(sc->stream);
    vb::DevBuf<int> d_idx(sc->stream);
    VB_CUDA(d_q.alloc(3 * (size_t)Q));
    VB_CUDA(d_d2.alloc((size_t)Q));
    VB_CUDA(d_idx.alloc((size_t)Q));
    VB_CUDA(cudaMemcpyAsync(d_q.p, q_xyz, sizeof(double) * 3 * (size_t)Q, cudaMemcpyHostToDevice, sc->stream));
    VB_TRY(vb::knn1_launch(sc, d_q.p, Q, radius, d_idx.p, d_d2.p));
    VB_CUDA(cudaMemcpyAsync(out_idx, d_idx.p, sizeof(int) * (size_t)Q, cudaMemcpyDeviceToHost, sc->stream));
    VB_CUDA(cudaMemcpyAsync(out_d2, d_d2.p, sizeof(double) * (size_t)Q, cudaMemcpyDeviceToHost, sc->stream));
    VB_CUDA(cudaStreamSynchronize(sc->stream));
    return VB200_OK;
}

extern "C" int vb200_knn1_bruteforce_device(const void *d_tgt_xyz, int64_t n, const void *d_q_xyz, int64_t Q,
                                            double radius, int device, void *d_out_idx, void *d_out_d2,
                                            void *cuda_stream) {
    if (n < 0 || Q < 0 || (n > 0 && !d_tgt_xyz) || (Q > 0 && (!d_q_xyz || !d_out_idx || !d_out_d2))) return VB200_ERR_INVALID;
    VB_TRY(vb::select_device(device));
    return vb::bf_launch((const double *)d_tgt_xyz, n, (const double *)d_q_xyz, Q, radius, (int *)d_out_idx,
                         (double *)d_out_d2, (cudaStream_t)cuda_stream);
}

extern "C" int vb200_knn1_bruteforce(const double *tgt_xyz, int64_t n, const double *q_xyz, int64_t Q, double radius,
                                     int device, int32_t *out_idx, double *out_d2) {
    if (n < 0 || Q < 0 || (n > 0 && !tgt_xyz) || (Q > 0 && (!q_xyz || !out_idx || !out_d2))) return VB200_ERR_INVALID;
    if (!(radius > 0.0)) return VB200_ERR_INVALID;
    VB_TRY(vb::select_device(device));
    if (Q == 0) return VB200_OK;
    cudaStream_t st;
    VB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};
    vb::DevBuf<double> d_t(st), d_q(st), d_d2(st);
    vb::DevBuf<int> d_idx(st);
    VB_CUDA(d_t.alloc(3 * (size_t)std::max<int64_t>(n, 1)));
    VB_CUDA(d_q.alloc(3 * (size_t)Q));
    VB_CUDA(d_d2.alloc((size_t)Q));
    VB_CUDA(d_idx.alloc((size_t)Q));
    if (n) VB_CUDA(vb::h2d_async(d_t.p, tgt_xyz, sizeof(double) * 3 * (size_t)n, st));
    VB_CUDA(cudaMemcpyAsync(d_q.p, q_xyz, sizeof(double) * 3 * (size_t)Q, cudaMemcpyHostToDevice, st));
    VB_TRY(vb::bf_launch(d_t.p, n, d_q.p, Q, radius, d_idx.p, d_d2.p, st));
    VB_CUDA(cudaMemcpyAsync(out_idx, d_idx.p, sizeof(int) * (size_t)Q, cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaMemcpyAsync(out_d2, d_d2.p, sizeof(double) * (size_t)Q, cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    return VB200_OK;
}
