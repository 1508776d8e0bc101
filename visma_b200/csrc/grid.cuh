// grid.cuh — the scene's nearest-neighbour structure and the device-side search that replaces the
// reference's FLANN KD-tree descent (O3D/src/Core/Geometry/KDTreeFlann.cpp:165-189 ->
// O3D/3rdparty/flann/algorithms/kdtree_single_index.h:593-642).
//
// Layout in HBM (all arrays in "sorted order": by coarse cell, then fine cell, then original index):
//   hi   float4[N]  (x-ctr, y-ctr, z-ctr) rounded to f32, w = original index bits   16 B/pt  screening
//   xyz  double[4N] exact coordinates, padded to one 32-byte sector                32 B/pt  decisions
//   nrm  double[4N] exact normals (optional), likewise                             32 B/pt  estimator
//   orig int[N]     sorted position -> original index                               4 B/pt
//   coarse CoarseCell[ncoarse]  {64-bit occupancy mask of the 4x4x4 fine cells, base into fstart} 16 B
//   fstart int[nfine+1]         start position of every OCCUPIED fine cell
// Coarse cell edge = max query radius, fine cell edge = 1/4 of it; a radius query touches at most 3x3x3
// coarse cells and each is one 16-byte load + a mask test.
//
// Exactness: candidates are screened with f32 distances on `hi` (coalescable float4 loads); every
// decision that could differ from the reference's double arithmetic — the winner among near-ties and the
// strict `d2 < (double)(float)(r*r)` acceptance — is re-taken in double on `xyz` using a rigorous error
// band on the f32 distance (see band()).
#pragma once

#include <math.h>

#include "common.cuh"

namespace vb {

struct __align__(16) CoarseCell {
    unsigned long long mask;  // bit (fx + 4 fy + 16 fz) set iff that fine cell holds points
    int base;                 // index into fstart of this coarse cell's first occupied fine cell
    int pad;
};

struct GridParams {
    double lo[3];       // corner of fine cell (0,0,0)
    double ctr[3];      // subtracted before rounding coordinates to f32
    double inv_fine;    // 1 / fine cell edge
    double cell;        // coarse cell edge (>= max radius)
    int fdim[3];        // grid size in fine cells (= 4 * coarse dims)
    int cdim[3];        // grid size in coarse cells
    float fine;         // fine cell edge, f32
    // |d32 - d2| <= band_a*sqrt(d32) + band_b + band_rel*d32 for every candidate/query pair in the grid
    float band_a, band_b, band_rel;
};

// Doubles between consecutive scene points in `xyz` and in `nrm`: records are padded to 32 bytes, one aligned L2
// sector each, and read with ONE 256-bit load (ld_rec).  With packed 24-byte records a gather cost three 8-byte
// requests of up to 32 sectors each and half the records straddled two sectors: the L1 tag stage, not DRAM, was
// what a settled correspondence pass saturated (l1tex throughput 63-74 % at 30 % of DRAM bandwidth).
constexpr int kPtStride = 4;

struct GridDev {
    GridParams p;
    const CoarseCell *coarse;
    const int *fstart;
    const float4 *hi;
    const double *xyz;  // kPtStride doubles per point (x, y, z, 0)
    const double *nrm;  // nullable; likewise
    const int *orig;
    int64_t n;
};

// f32 upper bound of the acceptance threshold r2 including the screening band (host side)
inline float r2_upper_bound(const GridParams &g, double r2) {
    float f = (float)(r2 * (1.0 + 1e-6));
    float bnd = g.band_a * sqrtf(f) + g.band_b + g.band_rel * f;
    f = f + 1.5f * bnd;
    return nextafterf(f, 3.0e38f);
}

#ifdef __CUDACC__

__device__ __forceinline__ float band(const GridParams &g, float d32) {
    return fmaf(g.band_a, sqrtf(d32), fmaf(g.band_rel, d32, g.band_b));
}

// flann::L2<double> on 3-D data: ((dx*dx) + dy*dy) + dz*dz with every operation rounded separately
// (O3D/3rdparty/flann/algorithms/dist.h:150-177) — no FMA contraction so d2 is bit-identical to the CPU.
// one scene record (x, y, z of a point or of a normal): a single 256-bit read-only load of its 32-byte sector
struct Rec3 {
    double x, y, z;
};
__device__ __forceinline__ Rec3 ld_rec(const double *__restrict__ rec) {
    Rec3 r;
    asm("{\n.reg .f64 pad;\nld.global.nc.v4.f64 {%0, %1, %2, pad}, [%3];\n}" : "=d"(r.x), "=d"(r.y), "=d"(r.z) : "l"(rec));
    return r;
}

// t = a scene record (G.xyz + kPtStride * position)
__device__ __forceinline__ double l2_exact(double qx, double qy, double qz, const double *__restrict__ t) {
    const Rec3 p = ld_rec(t);
    double dx = __dsub_rn(qx, p.x), dy = __dsub_rn(qy, p.y), dz = __dsub_rn(qz, p.z);
    double r = __dmul_rn(dx, dx);
    r = __dadd_rn(r, __dmul_rn(dy, dy));
    r = __dadd_rn(r, __dmul_rn(dz, dz));
    return r;
}

__device__ __forceinline__ double l2_exact(double qx, double qy, double qz, double tx, double ty, double tz) {
    double dx = __dsub_rn(qx, tx), dy = __dsub_rn(qy, ty), dz = __dsub_rn(qz, tz);
    double r = __dmul_rn(dx, dx);
    r = __dadd_rn(r, __dmul_rn(dy, dy));
    r = __dadd_rn(r, __dmul_rn(dz, dz));
    return r;
}

struct QueryCtx {
    float qx, qy, qz;  // centred f32 query
    float fx, fy, fz;  // position inside the home fine cell, in fine-cell units [0,1)
    int gx, gy, gz;    // home fine cell
};

// Returns false when the query is farther than one coarse cell outside the grid (cannot match).
__device__ __forceinline__ bool make_query(const GridParams &g, double x, double y, double z, QueryCtx &c) {
    double ux = (x - g.lo[0]) * g.inv_fine, uy = (y - g.lo[1]) * g.inv_fine, uz = (z - g.lo[2]) * g.inv_fine;
    // the negated comparisons also reject NaN
    if (!(ux >= 0.0 && uy >= 0.0 && uz >= 0.0 && ux < (double)g.fdim[0] && uy < (double)g.fdim[1] &&
          uz < (double)g.fdim[2]))
        return false;
    double fx = floor(ux), fy = floor(uy), fz = floor(uz);
    c.gx = (int)fx; c.gy = (int)fy; c.gz = (int)fz;
    c.fx = (float)(ux - fx); c.fy = (float)(uy - fy); c.fz = (float)(uz - fz);
    c.qx = (float)(x - g.ctr[0]); c.qy = (float)(y - g.ctr[1]); c.qz = (float)(z - g.ctr[2]);
    return true;
}

// make_query's range test alone
__device__ __forceinline__ bool inside_grid(const GridParams &g, double x, double y, double z) {
    const double ux = (x - g.lo[0]) * g.inv_fine, uy = (y - g.lo[1]) * g.inv_fine, uz = (z - g.lo[2]) * g.inv_fine;
    return ux >= 0.0 && uy >= 0.0 && uz >= 0.0 && ux < (double)g.fdim[0] && uy < (double)g.fdim[1] &&
           uz < (double)g.fdim[2];
}

// small non-negative int -> float without the (quarter-rate) I2F conversion unit
__device__ __forceinline__ float small_int_to_float(int v) { return __int_as_float(0x4B000000 | v) - 8388608.0f; }

// squared distance (fine-cell units) from the query to fine cell at integer offset o along one axis
__device__ __forceinline__ float axis_gap(int o, float frac) {
    float a = o > 0 ? (float)o - frac : (o < 0 ? frac - (float)(o + 1) : 0.0f);
    return a;
}

__device__ __forceinline__ unsigned long long range_mask(int ax, int bx, int ay, int by, int az, int bz) {
    // bits with fx in [ax,bx], fy in [ay,by], fz in [az,bz]; all in 0..3, a <= b
    unsigned long long xm = (unsigned long long)((1u << (bx + 1)) - (1u << ax)) * 0x1111111111111111ull;
    unsigned long long ym = (unsigned long long)((1u << (4 * (by + 1))) - (1u << (4 * ay))) * 0x0001000100010001ull;
    unsigned long long zhi = bz == 3 ? ~0ull : ((1ull << (16 * (bz + 1))) - 1ull);
    unsigned long long zm = zhi & ~((1ull << (16 * az)) - 1ull);
    return xm & ym & zm;
}

// ---- f32 screening walk ------------------------------------------------------------------------------
// On return: bs = sorted position of the f32-nearest candidate (or -1), best = its f32 d2,
// second = f32 d2 of the runner-up (or the initial bound).  `bound` = f32 upper bound of the acceptance
// threshold including its band.
struct Screen {
    float best, second;
    int bs;
};

__device__ __forceinline__ void scan_run(const float4 *__restrict__ hi, int s0, int s1, const QueryCtx &c,
                                         Screen &r) {
    auto consider = [&](const float4 t, int s) {
        const float dx = c.qx - t.x, dy = c.qy - t.y, dz = c.qz - t.z;
        const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        const bool lt = d < r.best;
        r.second = lt ? r.best : fminf(r.second, d);  // equal distances land in `second`: flagged ambiguous
        r.bs = lt ? s : r.bs;
        r.best = fminf(r.best, d);
    };
    // two candidates are ordered against each other first (independent of the running state), then merged:
    // the dependent chain through best / second / bs is one merge per pair instead of one per candidate
    auto consider2 = [&](const float4 ta, const float4 tb, int s) {
        const float ax = c.qx - ta.x, ay = c.qy - ta.y, az = c.qz - ta.z;
        const float bx = c.qx - tb.x, by = c.qy - tb.y, bz = c.qz - tb.z;
        const float da = fmaf(az, az, fmaf(ay, ay, ax * ax)), db = fmaf(bz, bz, fmaf(by, by, bx * bx));
        const float lo = fminf(da, db), hi2 = fmaxf(da, db);
        const int slo = db < da ? s + 1 : s;
        const bool lt = lo < r.best;
        r.second = fminf(fminf(r.second, hi2), fmaxf(lo, r.best));  // ties (lo == best, da == db) stay visible
        r.bs = lt ? slo : r.bs;
        r.best = fminf(r.best, lo);
    };
    int s = s0;
    for (; s + 3 < s1; s += 4) {  // four independent 16-byte loads in flight per iteration
        const float4 t0 = __ldg(hi + s), t1 = __ldg(hi + s + 1), t2 = __ldg(hi + s + 2), t3 = __ldg(hi + s + 3);
        consider2(t0, t1, s);
        consider2(t2, t3, s + 2);
    }
    const int rem = s1 - s;  // 0..3 left: issue their loads together, no loop
    if (rem > 0) {
        const float4 t0 = __ldg(hi + s);
        float4 t1 = t0, t2 = t0;
        if (rem > 1) t1 = __ldg(hi + s + 1);
        if (rem > 2) t2 = __ldg(hi + s + 2);
        consider(t0, s);
        if (rem > 1) consider(t1, s + 1);
        if (rem > 2) consider(t2, s + 2);
    }
}

template <class Visit>
__device__ __forceinline__ void walk_cells(const GridDev &G, const QueryCtx &c, float reach2_metric,
                                           bool skip_home, Visit &&visit) {
    // `visit(s0, s1, gap2_metric)` is called for every occupied fine cell whose box lies within
    // sqrt(reach2_metric) of the query at call time; the visitor re-checks against its current best.
    const GridParams &g = G.p;
    const float fine = g.fine;
    // reach in fine-cell units, padded: never prune a cell that could hold a closer point
    float rho = sqrtf(reach2_metric) / fine * 1.00001f + 1e-6f;
    int lx = -(int)ceilf(fmaxf(rho - c.fx, 0.0f)), hx = (int)ceilf(fmaxf(rho - (1.0f - c.fx), 0.0f));
    int ly = -(int)ceilf(fmaxf(rho - c.fy, 0.0f)), hy = (int)ceilf(fmaxf(rho - (1.0f - c.fy), 0.0f));
    int lz = -(int)ceilf(fmaxf(rho - c.fz, 0.0f)), hz = (int)ceilf(fmaxf(rho - (1.0f - c.fz), 0.0f));
    int fx0 = max(c.gx + lx, 0), fx1 = min(c.gx + hx, g.fdim[0] - 1);
    int fy0 = max(c.gy + ly, 0), fy1 = min(c.gy + hy, g.fdim[1] - 1);
    int fz0 = max(c.gz + lz, 0), fz1 = min(c.gz + hz, g.fdim[2] - 1);
    const int hcx = c.gx >> 2, hcy = c.gy >> 2, hcz = c.gz >> 2;
    const int hbit = (c.gx & 3) + 4 * (c.gy & 3) + 16 * (c.gz & 3);
    const float fine2 = fine * fine;
    for (int cz = fz0 >> 2; cz <= (fz1 >> 2); ++cz)
        for (int cy = fy0 >> 2; cy <= (fy1 >> 2); ++cy)
            for (int cx = fx0 >> 2; cx <= (fx1 >> 2); ++cx) {
                const CoarseCell cc = G.coarse[((int64_t)cz * g.cdim[1] + cy) * g.cdim[0] + cx];
                unsigned long long m = cc.mask;
                if (m == 0ull) continue;
                int ax = max(fx0 - 4 * cx, 0), bx = min(fx1 - 4 * cx, 3);
                int ay = max(fy0 - 4 * cy, 0), by = min(fy1 - 4 * cy, 3);
                int az = max(fz0 - 4 * cz, 0), bz = min(fz1 - 4 * cz, 3);
                unsigned long long sel = m & range_mask(ax, bx, ay, by, az, bz);
                if (skip_home && cx == hcx && cy == hcy && cz == hcz) sel &= ~(1ull << hbit);
                if (sel == 0ull) continue;
                // query position relative to this coarse cell's corner, in fine-cell units (conversions once
                // per coarse cell; the per-fine-cell offsets below avoid the quarter-rate I2F unit)
                const float rx = (float)(c.gx - 4 * cx) + c.fx, ry = (float)(c.gy - 4 * cy) + c.fy,
                            rz = (float)(c.gz - 4 * cz) + c.fz;
                while (sel) {
                    int b = __ffsll((long long)sel) - 1;
                    sel &= sel - 1ull;
                    const float fx = small_int_to_float(b & 3), fy = small_int_to_float((b >> 2) & 3),
                                fz = small_int_to_float(b >> 4);
                    const float gx = fmaxf(fmaxf(fx - rx, rx - fx - 1.0f), 0.0f);
                    const float gy = fmaxf(fmaxf(fy - ry, ry - fy - 1.0f), 0.0f);
                    const float gz = fmaxf(fmaxf(fz - rz, rz - fz - 1.0f), 0.0f);
                    // lower bound of the squared distance to anything in that cell (metric), deflated
                    float gap2 = (gx * gx + gy * gy + gz * gz) * fine2 * 0.998f - 1e-12f * fine2;
                    int rank = __popcll(m & ((1ull << b) - 1ull));
                    int s0 = __ldg(G.fstart + cc.base + rank), s1 = __ldg(G.fstart + cc.base + rank + 1);
                    visit(s0, s1, gap2);
                }
            }
}

// Full radius-bounded 1-NN for one query.  r2 = (double)(float)(radius*radius), r2_ub = f32 bound
// >= r2 + band.  Returns sorted position of the accepted neighbour or -1; d2 = exact double distance.
static __device__ __noinline__ int nn_exact_rescan(const GridDev &G, const QueryCtx &c, double qx, double qy,
                                            double qz, double r2, float reach2, double *d2_out) {
    // Rare path (near-tie between candidates, or winner within the band of the threshold): redo the walk
    // in double.  Ties break to the lowest ORIGINAL index (documented rule; FLANN's is traversal order).
    double bd = r2;
    int bs = -1, bo = 0x7fffffff;
    auto visit = [&](int s0, int s1, float gap2) {
        if ((double)gap2 > bd) return;
        for (int s = s0; s < s1; ++s) {
            double d = l2_exact(qx, qy, qz, G.xyz + kPtStride * (int64_t)s);
            if (d < bd) {
                bd = d; bs = s; bo = __ldg(G.orig + s);
            } else if (d == bd && bs >= 0) {
                int o = __ldg(G.orig + s);
                if (o < bo) { bs = s; bo = o; }
            }
        }
    };
    walk_cells(G, c, reach2, false, visit);
    *d2_out = bs >= 0 ? bd : 0.0;
    return bs;
}

// ---- warp-cooperative search -----------------------------------------------------------------------------
// A per-lane cell walk diverges badly when the 32 lanes of a warp visit different cells (first version,
// measured: 4.5 of 32 lanes active per issued instruction).  Instead the warp repeatedly picks a leader lane,
// groups the pending lanes whose home cell is within +-2 fine cells of the leader's, and walks ONE box of fine
// cells that covers every group member's reach: all loop bounds, the coarse-cell loads, the fine-cell ranges
// and the candidate loads are warp-uniform (one broadcast load serves the whole group) and each lane only
// keeps its own best / runner-up.  Scanning extra candidates never changes a lane's result, so the answers
// are those of an independent per-query search.  k_pass sorts every source cloud spatially so that a warp is
// normally one or two groups.

struct FineBox { int x0, x1, y0, y1, z0, z1; };

__device__ __forceinline__ bool box_empty(const FineBox &b) { return b.x0 > b.x1 || b.y0 > b.y1 || b.z0 > b.z1; }


// upper bound of best + 2*band(best), capped at the acceptance bound (fast reciprocal sqrt, padded)
__device__ __forceinline__ float reach_of(const GridParams &g, float best, float r2_ub) {
    const float b = fmaxf(best, 1e-30f);
    const float bnd = fmaf(g.band_a * 1.0001f, b * rsqrtf(b), fmaf(g.band_rel, b, g.band_b));
    return fminf(fmaf(2.0f, bnd, best), r2_ub);
}

// lane position in absolute fine-cell units (cell index + fraction); resolution 2^-12 cell or better, which
// the 1e-3 relative deflation of the gap below absorbs
struct LanePos { float ux, uy, uz; };

// scan every occupied fine cell of `box` (minus `skip`, if given) with warp-uniform control flow
template <bool PRUNE, bool HAS_SKIP>
__device__ __forceinline__ void scan_box_uniform(const GridDev &G, const QueryCtx &c, const LanePos &lp, bool member,
                                                 const FineBox &box, const FineBox &skip, float r2_ub, Screen &r) {
    const GridParams &g = G.p;
    const float fine2 = g.fine * g.fine * 0.998f;  // deflated: never prune a cell that could matter
    float thr = reach_of(g, r.best, r2_ub);
    for (int cz = box.z0 >> 2; cz <= (box.z1 >> 2); ++cz)
        for (int cy = box.y0 >> 2; cy <= (box.y1 >> 2); ++cy)
            for (int cx = box.x0 >> 2; cx <= (box.x1 >> 2); ++cx) {
                const CoarseCell cc = G.coarse[((int64_t)cz * g.cdim[1] + cy) * g.cdim[0] + cx];
                const unsigned long long m = cc.mask;
                if (m == 0ull) continue;
                unsigned long long sel = m & range_mask(max(box.x0 - 4 * cx, 0), min(box.x1 - 4 * cx, 3),
                                                        max(box.y0 - 4 * cy, 0), min(box.y1 - 4 * cy, 3),
                                                        max(box.z0 - 4 * cz, 0), min(box.z1 - 4 * cz, 3));
                if (HAS_SKIP) {
                    int ax = max(skip.x0 - 4 * cx, 0), bx = min(skip.x1 - 4 * cx, 3);
                    int ay = max(skip.y0 - 4 * cy, 0), by = min(skip.y1 - 4 * cy, 3);
                    int az = max(skip.z0 - 4 * cz, 0), bz = min(skip.z1 - 4 * cz, 3);
                    if (ax <= bx && ay <= by && az <= bz) sel &= ~range_mask(ax, bx, ay, by, az, bz);
                }
                if (sel == 0ull) continue;
                // lane position relative to this coarse cell's corner, fine-cell units
                const float rx = lp.ux - (float)(4 * cx), ry = lp.uy - (float)(4 * cy), rz = lp.uz - (float)(4 * cz);
                while (sel) {
                    const int b = __ffsll((long long)sel) - 1;
                    sel &= sel - 1ull;
                    if (PRUNE) {
                        // skip the cell only if NO member can have its winner (or an ambiguous runner-up) in it
                        const float fx = small_int_to_float(b & 3), fy = small_int_to_float((b >> 2) & 3),
                                    fz = small_int_to_float(b >> 4);
                        const float ex = fmaxf(fmaxf(fx - rx, rx - fx - 1.0f), 0.0f);
                        const float ey = fmaxf(fmaxf(fy - ry, ry - fy - 1.0f), 0.0f);
                        const float ez = fmaxf(fmaxf(fz - rz, rz - fz - 1.0f), 0.0f);
                        const float gap2 = (ex * ex + ey * ey + ez * ez) * fine2;
                        if (!__any_sync(0xffffffffu, member && gap2 <= thr)) continue;
                    }
                    const int rank = __popcll(m & ((1ull << b) - 1ull));
                    const int s0 = __ldg(G.fstart + cc.base + rank), s1 = __ldg(G.fstart + cc.base + rank + 1);
                    scan_run(G.hi, s0, s1, c, r);
                    if (PRUNE) thr = reach_of(g, r.best, r2_ub);
                }
            }
}

// Phase-2 scan in rounds.  Each active lane asks for the cells within `want` (squared metric reach) that lie
// beyond what it has already completed (`done`); a cell is scanned when some active lane asks for it and no
// member has it inside its completed region (then it was scanned in an earlier round).  Lanes that still have
// no candidate grow their reach one fine cell per round instead of jumping to the full radius — an empty home
// cell next to a populated surface would otherwise drag the whole warp through every cell of the radius ball.
// thr_end returns the lane's final threshold: every cell with gap2 <= thr_end has been scanned.
__device__ __forceinline__ void scan_box_annulus(const GridDev &G, const QueryCtx &c, const LanePos &lp, bool member,
                                                 bool active, const FineBox &box, const FineBox &skip, float want,
                                                 float done, float r2_ub, Screen &r, float &thr_end) {
    const GridParams &g = G.p;
    const float fine2 = g.fine * g.fine * 0.998f;  // deflated: never prune a cell that could matter
    float thr = want;
    for (int cz = box.z0 >> 2; cz <= (box.z1 >> 2); ++cz)
        for (int cy = box.y0 >> 2; cy <= (box.y1 >> 2); ++cy)
            for (int cx = box.x0 >> 2; cx <= (box.x1 >> 2); ++cx) {
                const CoarseCell cc = G.coarse[((int64_t)cz * g.cdim[1] + cy) * g.cdim[0] + cx];
                const unsigned long long m = cc.mask;
                if (m == 0ull) continue;
                unsigned long long sel = m & range_mask(max(box.x0 - 4 * cx, 0), min(box.x1 - 4 * cx, 3),
                                                        max(box.y0 - 4 * cy, 0), min(box.y1 - 4 * cy, 3),
                                                        max(box.z0 - 4 * cz, 0), min(box.z1 - 4 * cz, 3));
                {
                    int ax = max(skip.x0 - 4 * cx, 0), bx = min(skip.x1 - 4 * cx, 3);
                    int ay = max(skip.y0 - 4 * cy, 0), by = min(skip.y1 - 4 * cy, 3);
                    int az = max(skip.z0 - 4 * cz, 0), bz = min(skip.z1 - 4 * cz, 3);
                    if (ax <= bx && ay <= by && az <= bz) sel &= ~range_mask(ax, bx, ay, by, az, bz);
                }
                if (sel == 0ull) continue;
                const float rx = lp.ux - (float)(4 * cx), ry = lp.uy - (float)(4 * cy), rz = lp.uz - (float)(4 * cz);
                while (sel) {
                    const int b = __ffsll((long long)sel) - 1;
                    sel &= sel - 1ull;
                    const float fx = small_int_to_float(b & 3), fy = small_int_to_float((b >> 2) & 3),
                                fz = small_int_to_float(b >> 4);
                    const float ex = fmaxf(fmaxf(fx - rx, rx - fx - 1.0f), 0.0f);
                    const float ey = fmaxf(fmaxf(fy - ry, ry - fy - 1.0f), 0.0f);
                    const float ez = fmaxf(fmaxf(fz - rz, rz - fz - 1.0f), 0.0f);
                    const float gap2 = (ex * ex + ey * ey + ez * ez) * fine2;
                    const bool asks = active && gap2 <= thr && gap2 > done;
                    const bool had = member && gap2 <= done;
                    if (!__any_sync(0xffffffffu, asks) || __any_sync(0xffffffffu, had)) continue;
                    const int rank = __popcll(m & ((1ull << b) - 1ull));
                    const int s0 = __ldg(G.fstart + cc.base + rank), s1 = __ldg(G.fstart + cc.base + rank + 1);
                    scan_run(G.hi, s0, s1, c, r);
                    if (r.bs >= 0) thr = fminf(thr, reach_of(g, r.best, r2_ub));
                }
            }
    thr_end = thr;
}

#ifndef VB_GROUP_HW
#define VB_GROUP_HW 2
#endif
constexpr int kGroupHalfWidth = VB_GROUP_HW;  // lanes within +-2 fine cells of the leader are searched together

// f32 screening for the lanes of `pending` (warp-uniform mask), searched group by group.  All 32 lanes of
// the warp must call this together; lanes outside `pending` ride along and get an empty Screen back.
__device__ __forceinline__ Screen coop_screen(const GridDev &G, unsigned pending, const QueryCtx &c, float r2_ub) {
    const unsigned FULL = 0xffffffffu;
    const GridParams &g = G.p;
    Screen r;
    r.best = r2_ub; r.second = 3.0e38f; r.bs = -1;
    LanePos lp;
    lp.ux = (float)c.gx + c.fx; lp.uy = (float)c.gy + c.fy; lp.uz = (float)c.gz + c.fz;
    const int lane = threadIdx.x & 31;
    while (pending) {
        const int leader = __ffs(pending) - 1;
        const int lgx = __shfl_sync(FULL, c.gx, leader), lgy = __shfl_sync(FULL, c.gy, leader),
                  lgz = __shfl_sync(FULL, c.gz, leader);
        const bool member = ((pending >> lane) & 1u) && abs(c.gx - lgx) <= kGroupHalfWidth &&
                            abs(c.gy - lgy) <= kGroupHalfWidth && abs(c.gz - lgz) <= kGroupHalfWidth;
        pending &= ~__ballot_sync(FULL, member);
        QueryCtx cq = c;
        Screen rr;
        rr.best = r2_ub; rr.second = 3.0e38f; rr.bs = -1;
        if (!member) { cq.qx = 3.0e18f; cq.qy = 3.0e18f; cq.qz = 3.0e18f; }  // never the best of anything
        // phase 1: the cells that contain the group's queries — every member gets a tight bound from its own cell
        FineBox b0;
        b0.x0 = __reduce_min_sync(FULL, member ? c.gx : 0x7fffffff); b0.x1 = __reduce_max_sync(FULL, member ? c.gx : -0x7fffffff);
        b0.y0 = __reduce_min_sync(FULL, member ? c.gy : 0x7fffffff); b0.y1 = __reduce_max_sync(FULL, member ? c.gy : -0x7fffffff);
        b0.z0 = __reduce_min_sync(FULL, member ? c.gz : 0x7fffffff); b0.z1 = __reduce_max_sync(FULL, member ? c.gz : -0x7fffffff);
        scan_box_uniform<false, false>(G, cq, lp, member, b0, b0, r2_ub, rr);
        // phase 2, in rounds: members with a candidate ask for their exact reach (one round); members without
        // one grow their reach a fine cell per round until something turns up or the radius is exhausted
        bool fin = !member;
        float done = -1.0f;
        for (int round = 1;; ++round) {
            const bool active = member && !fin;
            if (!__any_sync(FULL, active)) break;
            const float ringr = (float)round * g.fine;
            const float want = rr.bs >= 0 ? reach_of(g, rr.best, r2_ub) : fminf(ringr * ringr, r2_ub);
            const float rho = sqrtf(want) / g.fine * 1.001f + 1e-4f;
            FineBox need;
            need.x0 = c.gx - (int)ceilf(fmaxf(rho - c.fx, 0.0f)); need.x1 = c.gx + (int)ceilf(fmaxf(rho - (1.0f - c.fx), 0.0f));
            need.y0 = c.gy - (int)ceilf(fmaxf(rho - c.fy, 0.0f)); need.y1 = c.gy + (int)ceilf(fmaxf(rho - (1.0f - c.fy), 0.0f));
            need.z0 = c.gz - (int)ceilf(fmaxf(rho - c.fz, 0.0f)); need.z1 = c.gz + (int)ceilf(fmaxf(rho - (1.0f - c.fz), 0.0f));
            FineBox b1;
            b1.x0 = __reduce_min_sync(FULL, active ? need.x0 : 0x7fffffff); b1.x1 = __reduce_max_sync(FULL, active ? need.x1 : -0x7fffffff);
            b1.y0 = __reduce_min_sync(FULL, active ? need.y0 : 0x7fffffff); b1.y1 = __reduce_max_sync(FULL, active ? need.y1 : -0x7fffffff);
            b1.z0 = __reduce_min_sync(FULL, active ? need.z0 : 0x7fffffff); b1.z1 = __reduce_max_sync(FULL, active ? need.z1 : -0x7fffffff);
            b1.x0 = max(b1.x0, 0); b1.y0 = max(b1.y0, 0); b1.z0 = max(b1.z0, 0);
            b1.x1 = min(b1.x1, g.fdim[0] - 1); b1.y1 = min(b1.y1, g.fdim[1] - 1); b1.z1 = min(b1.z1, g.fdim[2] - 1);
            float thr_end = want;
            if (!box_empty(b1)) scan_box_annulus(G, cq, lp, member, active, b1, b0, want, done, r2_ub, rr, thr_end);
            if (active) {
                // complete once the need computed from the final best fits inside what this round covered
                fin = want >= r2_ub || (rr.bs >= 0 && reach_of(g, rr.best, r2_ub) <= want);
                done = thr_end;
            }
        }
        if (member) r = rr;
    }
    return r;
}

// ---- lane-private search ----------------------------------------------------------------------------------
// The cooperative walk above makes every lane of a group evaluate every candidate ANY member needs: ~100-170
// candidates per 32 queries where each query alone needs 10-30.  Once a lane has a bound on its answer — the
// distance to its match of the previous ICP iteration (`prior`), or the best candidate of its own home cell —
// its reach is a fraction of a fine cell and the cells it needs are few.  Such a lane lists ITS cells (runs of
// the sorted arrays) in a private column of shared memory and scans only those in a flat loop: neighbouring
// lanes walk the same runs in step, so the loads still coalesce to a handful of 16-byte segments, but the
// warp's trip count is the LARGEST lane's candidate count instead of the union's.  Lanes without a bound, with
// a reach beyond kLaneMaxRho fine cells, or with more than kLaneMaxRuns cells fall back to coop_screen.  The
// screening rule (every candidate within best + 2*band scanned, ambiguity flagged) is the same, so results are
// those of the other searches, bit for bit.
#ifndef VB_LANE_MAX_RUNS
#define VB_LANE_MAX_RUNS 16
#endif
#ifndef VB_LANE_MAX_RHO_PCT
#define VB_LANE_MAX_RHO_PCT 125
#endif
constexpr int kLaneMaxRuns = VB_LANE_MAX_RUNS;
constexpr float kLaneMaxRho = VB_LANE_MAX_RHO_PCT * 0.01f;  // largest reach (fine-cell units) searched per lane
#ifndef VB_LANE_PROBE_RHO_PCT
#define VB_LANE_PROBE_RHO_PCT 50
#endif
constexpr float kLaneProbeRho = VB_LANE_PROBE_RHO_PCT * 0.01f;  // first reach of a lane that has no bound yet
#ifndef VB_BIG_RHO_PCT
#define VB_BIG_RHO_PCT 75
#endif
constexpr float kLaneBigRho = VB_BIG_RHO_PCT * 0.01f;  // a reach beyond this counts towards sending the warp to the shared walk
constexpr unsigned kRunLenBits = 12, kRunLenMask = (1u << kRunLenBits) - 1u;
#ifndef VB_NB_WINDOW
#define VB_NB_WINDOW 2
#endif
constexpr int kNbWindow = VB_NB_WINDOW;  // previous matches of the lanes within +-this many give the search its bound

#ifdef VB_STATS
__device__ unsigned long long g_stats[24];
#define VB_STAT(i, v) atomicAdd(&g_stats[i], (unsigned long long)(v))
#else
#define VB_STAT(i, v) ((void)0)
#endif

// (sqrt(thr) + slack)^2, capped — never below thr itself
__device__ __forceinline__ float widen(float thr, float slack, float cap) {
    const float s = sqrtf(thr) + slack;
    return fmaxf(thr, fminf(s * s, cap));
}

#ifndef VB_LANE_MAX_NEAR
#define VB_LANE_MAX_NEAR 8
#endif
constexpr int kLaneMaxNear = VB_LANE_MAX_NEAR;

template <int TPB>
struct LaneRuns {  // one column per thread: conflict-free whatever row each lane is at
    int s0[kLaneMaxRuns][TPB];
    unsigned w[kLaneMaxRuns][TPB];  // drop threshold h (f32, low 12 mantissa bits dropped = rounded down) | run length
    int near[kLaneMaxNear][TPB];    // settling searches: positions of the candidates nearer than the listing threshold
};

// scan this lane's listed runs; `bound` = f32 distance of a real candidate (or >= r2_ub): cells beyond its reach
// cannot matter.  A run carries h, computed when it was listed: while min(bound, best so far) < h the run lies
// beyond the (widened) reach of the best so far and is dropped (see list_runs_union) — one compare per run.
// SET: the positions of the candidates with d < near_thr are noted in the lane's column (nnear counts them all).
template <int TPB, bool SET>
__device__ __forceinline__ int scan_runs(const GridDev &G, const QueryCtx &c, int nruns, float bound, LaneRuns<TPB> &L,
                                         Screen &r, float near_thr, int &nnear) {
    const float4 *__restrict__ hi = G.hi;
    const int tid = threadIdx.x & (TPB - 1);
    int ri = 0, s = 0, e = 0;
    int steps = 0;  // dev statistics only; dead code otherwise
    for (;;) {
        if (s >= e) {  // next listed run (lanes that share cells get here together)
            if (ri >= nruns) break;
            const unsigned w = L.w[ri][tid];
            s = L.s0[ri][tid];
            e = s + (int)(w & kRunLenMask);
            ++ri;
            if (fminf(bound, r.best) < __uint_as_float(w & ~kRunLenMask)) e = s;
            continue;
        }
        // two candidates per step; the second is masked out when the run has one left
        const bool two = s + 1 < e;
        const float4 ta = __ldg(hi + s), tb = __ldg(hi + (two ? s + 1 : s));
        const float ax = c.qx - ta.x, ay = c.qy - ta.y, az = c.qz - ta.z;
        const float bx = c.qx - tb.x, by = c.qy - tb.y, bz = c.qz - tb.z;
        const float da = fmaf(az, az, fmaf(ay, ay, ax * ax));
        const float db = two ? fmaf(bz, bz, fmaf(by, by, bx * bx)) : 3.0e38f;
        if (SET) {
            if (da < near_thr) { if (nnear < kLaneMaxNear) L.near[nnear][tid] = s; ++nnear; }
            if (db < near_thr) { if (nnear < kLaneMaxNear) L.near[nnear][tid] = s + 1; ++nnear; }
        }
        const float lo = fminf(da, db), hi2 = fmaxf(da, db);
        const int slo = db < da ? s + 1 : s;
        const bool lt = lo < r.best;
        r.second = fminf(fminf(r.second, hi2), fmaxf(lo, r.best));
        r.bs = lt ? slo : r.bs;
        r.best = fminf(r.best, lo);
        s += 2;
        ++steps;
    }
    return steps;
}

// The five smallest f32 distances among the candidates a settling search noted (scan_runs<SET>) and the positions
// of the first four: the lane's candidate SET (see nn_search_hybrid).
struct Top5 {
    float e[5];
    int s[4];
};

template <int TPB>
__device__ __forceinline__ Top5 top5_of_near(const GridDev &G, const QueryCtx &c, int nnear, const LaneRuns<TPB> &L) {
    const int tid = threadIdx.x & (TPB - 1);
    float e0 = 3.0e38f, e1 = 3.0e38f, e2 = 3.0e38f, e3 = 3.0e38f, e4 = 3.0e38f;
    int s0 = -1, s1 = -1, s2 = -1, s3 = -1;
    for (int i = 0; i < nnear; ++i) {
        const int s = L.near[i][tid];
        const float4 p = __ldg(G.hi + s);
        const float dx = c.qx - p.x, dy = c.qy - p.y, dz = c.qz - p.z;
        const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        // sorted insertion (strict: equal distances keep their scan order)
        const bool l4 = d < e4, l3 = d < e3, l2 = d < e2, l1 = d < e1, l0 = d < e0;
        e4 = l3 ? e3 : (l4 ? d : e4);
        e3 = l2 ? e2 : (l3 ? d : e3); s3 = l2 ? s2 : (l3 ? s : s3);
        e2 = l1 ? e1 : (l2 ? d : e2); s2 = l1 ? s1 : (l2 ? s : s2);
        e1 = l0 ? e0 : (l1 ? d : e1); s1 = l0 ? s0 : (l1 ? s : s1);
        e0 = l0 ? d : e0;             s0 = l0 ? s : s0;
    }
    Top5 t;
    t.e[0] = e0; t.e[1] = e1; t.e[2] = e2; t.e[3] = e3; t.e[4] = e4;
    t.s[0] = s0; t.s[1] = s1; t.s[2] = s2; t.s[3] = s3;
    return t;
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// one listed run: its start / length and the drop threshold h (see list_runs)
template <int TPB>
__device__ __forceinline__ void store_run(const GridParams &g, LaneRuns<TPB> &L, int n, int tid, int s0, int s1, float gap2,
                                          float slack) {
    // (h only has to err low: the square root comes from the 2-ulp reciprocal-square-root unit, deflated; sqrt(gp) = sg)
    const float sg = fmaxf(gap2 * rsqrtf(fmaxf(gap2, 1e-30f)) * 0.99999f - slack, 0.0f) * 0.999999f;
    const float gp = sg * sg;
    const float bnd = fmaf(g.band_a * 1.001f, sg, fmaf(g.band_rel * 1.001f, gp, g.band_b * 1.001f));
    const float h = fmaxf(fmaf(-2.0f, bnd, gp), 0.0f);
    L.s0[n][tid] = s0;
    L.w[n][tid] = (__float_as_uint(h) & ~kRunLenMask) | (unsigned)(s1 - s0);
}

// list the occupied fine cells with done < gap2 <= thr (squared distance to the query, deflated) as runs, in ONE
// flat loop over the cells of the query's box (nested loops with per-lane trip counts serialise: measured 30 passes
// of the inner body per warp for 2.5 runs per lane; a warp-cooperative walk over the UNION of the lanes' boxes was
// measured twice as slow as this: the union of 32 neighbouring boxes holds ~5x the occupied cells of one box and
// every lane pays for each).  Returns the number of runs, or kLaneMaxRuns + 1 when the lane must fall back.
// Each listed run gets h: while min(bound, best so far) < h the scan may drop the run.  With W(b) =
// widen(reach_of(b), slack) the scan must keep every run with gap2 <= W(final best); h = g' - 2 bnd(g') with
// sqrt(g') = sqrt(gap2) - slack satisfies b < h  =>  (sqrt(reach_of(b)) + slack)^2 < gap2  =>  W(b) < gap2.
template <int TPB>
__device__ __forceinline__ int list_runs(const GridDev &G, const QueryCtx &c, float thr, float done, float slack,
                                         LaneRuns<TPB> &L) {
    const GridParams &g = G.p;
    const int tid = threadIdx.x & (TPB - 1);
    const float fine2 = g.fine * g.fine;
    // reach in fine cells.  The box must contain EVERY cell the deflated gap test below accepts (gap2 * 0.998
    // <= thr, i.e. up to 1.001 x the reach): a later round skips cells with gap2 <= done as already scanned.
    const float rho = sqrtf(thr) / g.fine * 1.0011f + 1e-4f;
    const int x0 = max(c.gx - (int)ceilf(fmaxf(rho - c.fx, 0.0f)), 0),
              x1 = min(c.gx + (int)ceilf(fmaxf(rho - (1.0f - c.fx), 0.0f)), g.fdim[0] - 1);
    const int y0 = max(c.gy - (int)ceilf(fmaxf(rho - c.fy, 0.0f)), 0),
              y1 = min(c.gy + (int)ceilf(fmaxf(rho - (1.0f - c.fy), 0.0f)), g.fdim[1] - 1);
    const int z0 = max(c.gz - (int)ceilf(fmaxf(rho - c.fz, 0.0f)), 0),
              z1 = min(c.gz + (int)ceilf(fmaxf(rho - (1.0f - c.fz), 0.0f)), g.fdim[2] - 1);
    int n = 0;
    int x = x0, y = y0, z = z0;
    // cell offsets from the home cell as floats, stepped with the integers (no conversions in the loop)
    const float ox0 = (float)(x0 - c.gx), oy0 = (float)(y0 - c.gy);
    float ox = ox0, oy = oy0, oz = (float)(z0 - c.gz);
    for (;;) {
        // squared distance from the query to cell (x, y, z), deflated: never drop a cell that could matter
        const float ex = fmaxf(fmaxf(ox - c.fx, c.fx - ox - 1.0f), 0.0f);
        const float ey = fmaxf(fmaxf(oy - c.fy, c.fy - oy - 1.0f), 0.0f);
        const float ez = fmaxf(fmaxf(oz - c.fz, c.fz - oz - 1.0f), 0.0f);
        const float gap2 = fmaxf((ex * ex + ey * ey + ez * ez) * fine2 * 0.998f - 1e-12f * fine2, 0.0f);
        if (gap2 <= thr && gap2 > done) {
            const CoarseCell cc = G.coarse[((z >> 2) * g.cdim[1] + (y >> 2)) * g.cdim[0] + (x >> 2)];
            const int bit = (x & 3) + 4 * (y & 3) + 16 * (z & 3);
            if ((cc.mask >> bit) & 1ull) {
                const int rank = __popcll(cc.mask & ((1ull << bit) - 1ull));
                const int s0 = __ldg(G.fstart + cc.base + rank), s1 = __ldg(G.fstart + cc.base + rank + 1);
                if (n >= kLaneMaxRuns || s1 - s0 > (int)kRunLenMask) {
                    n = kLaneMaxRuns + 1;  // fall back; the remaining cells are skipped by the test below
                    z = z1; y = y1; x = x1;
                } else {
                    prefetch_l1(G.hi + s0);
                    store_run<TPB>(g, L, n, tid, s0, s1, gap2, slack);
                    ++n;
                }
            }
        }
        // step (x, y, z) through the box with selects only: one backward branch, so the lanes of a warp stay
        // converged for max(box volume) trips
        const bool wx = x == x1, wy = wx && y == y1;
        x = wx ? x0 : x + 1;
        ox = wx ? ox0 : ox + 1.0f;
        y = wy ? y0 : (wx ? y + 1 : y);
        oy = wy ? oy0 : (wx ? oy + 1.0f : oy);
        z += wy ? 1 : 0;
        oz += wy ? 1.0f : 0.0f;
        if (z > z1) break;
    }
    return n;
}

// The same list by ROWS of the box instead of cells, for a lane's first round (nothing scanned yet).  Inside a coarse
// cell the fine cells are stored in the order x + 4 y + 16 z, so the cells of one box row (fixed y, z; consecutive x)
// that share a coarse cell are ONE contiguous run of the sorted arrays: its ends are two ranks in the coarse cell's
// occupancy mask.  A box of up to 4 x 4 x 4 cells is walked as at most 16 rows x 2 segments (a row crosses at most one
// coarse-cell boundary) instead of up to 64 cells — the walk over the cells of a lane's box was 38 % of part B's
// instructions at 11.7 of 32 lanes.  Each row is trimmed to the cells within reach along x (one square root); a
// run's drop threshold comes from the distance to its nearest cell.  The listed cells are a superset of those the
// per-cell test accepts, which is all the search's bookkeeping needs (every cell with gap2 <= thr is scanned).
template <int TPB>
__device__ __forceinline__ int list_rows(const GridDev &G, const QueryCtx &c, float thr, float slack, LaneRuns<TPB> &L) {
    const GridParams &g = G.p;
    const int tid = threadIdx.x & (TPB - 1);
    const float fine2 = g.fine * g.fine;
    const float rho = sqrtf(thr) / g.fine * 1.0011f + 1e-4f;  // (as list_runs: covers the deflated gap test)
    const int x0 = max(c.gx - (int)ceilf(fmaxf(rho - c.fx, 0.0f)), 0),
              x1 = min(c.gx + (int)ceilf(fmaxf(rho - (1.0f - c.fx), 0.0f)), g.fdim[0] - 1);
    const int y0 = max(c.gy - (int)ceilf(fmaxf(rho - c.fy, 0.0f)), 0),
              y1 = min(c.gy + (int)ceilf(fmaxf(rho - (1.0f - c.fy), 0.0f)), g.fdim[1] - 1);
    const int z0 = max(c.gz - (int)ceilf(fmaxf(rho - c.fz, 0.0f)), 0),
              z1 = min(c.gz + (int)ceilf(fmaxf(rho - (1.0f - c.fz), 0.0f)), g.fdim[2] - 1);
    if (x1 - x0 > 3) return kLaneMaxRuns + 1;  // (never at a lane-private reach: a row would span three coarse cells)
    // the two x segments: [x0, xm] in the first coarse cell, (xm, x1] in the next one (empty unless the row crosses)
    const int xm = min(x1, x0 | 3);
    const bool two = x1 > xm;
    // the reach in cell units, inflated like the box: a cell is within reach along x iff its ex <= sqrt(rho2 - ey^2 - ez^2)
    const float rho2 = rho * rho;
    int n = 0;
    int y = y0, z = z0;
    bool second = false;
    const float oy0 = (float)(y0 - c.gy);
    float oy = oy0, oz = (float)(z0 - c.gz);
    for (;;) {
        const float ey = fmaxf(fmaxf(oy - c.fy, c.fy - oy - 1.0f), 0.0f);
        const float ez = fmaxf(fmaxf(oz - c.fz, c.fz - oz - 1.0f), 0.0f);
        const float eyz2 = ey * ey + ez * ez;
        const float w2 = rho2 - eyz2;
        if (w2 >= 0.0f) {
            // cells of this segment within reach along x: offsets ox (from the home cell) with max(ox - fx, fx - ox - 1) <= w
            // (sqrt by the 2-ulp reciprocal-square-root unit, inflated: only a superset is needed; conversions with the
            // rounding built in)
            const float w = fmaf(w2, rsqrtf(fmaxf(w2, 1e-30f)), 1e-5f) * 1.0001f;
            const int sa = second ? xm + 1 : x0, sb = second ? x1 : xm;
            const int xa = max(sa, c.gx + __float2int_ru(c.fx - w - 1.0f)), xb = min(sb, c.gx + __float2int_rd(c.fx + w));
            if (xa <= xb) {
                const CoarseCell cc = G.coarse[((z >> 2) * g.cdim[1] + (y >> 2)) * g.cdim[0] + (xa >> 2)];
                const int b0 = (xa & 3) + 4 * (y & 3) + 16 * (z & 3), b1 = b0 + (xb - xa);
                const unsigned long long below0 = (1ull << b0) - 1ull, upto1 = b1 == 63 ? ~0ull : ((2ull << b1) - 1ull);
                const int r0 = __popcll(cc.mask & below0), r1 = __popcll(cc.mask & upto1);
                if (r1 > r0) {
                    const int s0 = __ldg(G.fstart + cc.base + r0), s1 = __ldg(G.fstart + cc.base + r1);
                    if (n >= kLaneMaxRuns || s1 - s0 > (int)kRunLenMask) {
                        n = kLaneMaxRuns + 1;  // fall back; end the walk
                        z = z1; y = y1; second = two;
                    } else {
                        prefetch_l1(G.hi + s0);
                        // distance to the segment's nearest cell, deflated as in list_runs
                        const float oxa = (float)(xa - c.gx), oxb = (float)(xb - c.gx);
                        const float ex = fmaxf(fmaxf(oxa - c.fx, c.fx - oxb - 1.0f), 0.0f);
                        const float gap2 = fmaxf((ex * ex + eyz2) * fine2 * 0.998f - 1e-12f * fine2, 0.0f);
                        store_run<TPB>(g, L, n, tid, s0, s1, gap2, slack);
                        ++n;
                    }
                }
            }
        }
        // step (segment, y, z) with selects only: one backward branch, the lanes of a warp stay converged
        const bool ws = second || !two, wy = ws && y == y1;
        second = ws ? false : true;
        y = wy ? y0 : (ws ? y + 1 : y);
        oy = wy ? oy0 : (ws ? oy + 1.0f : oy);
        z += wy ? 1 : 0;
        oz += wy ? 1.0f : 0.0f;
        if (z > z1) break;
    }
    return n;
}

// What a search proved about every target point it did NOT return, for the caller's cached-neighbour tests
// (k_pass_a: Greenspan & Godin's test for ICP, made exact).  All bounds are on TRUE squared distances from the
// query position of this search.
struct SearchProof {
    float sec1;     // > 0: every target point other than the returned one is at least this far; <= 0: nothing known
    float secK;     // > 0: every target point other than the returned one and others[] is at least this far
    int others[3];  // the rest of the candidate set (sorted positions, -1 = unused), nearest first
};

// All 32 lanes of the warp must call this together.  `prior` = sorted position of a target point believed to
// be close to the query (or -1): only ever used as an upper bound, never as an answer.
//
// `slack` (metric, >= 0) widens every lane-private search beyond what the answer needs; with `want_set` a lane
// that completed its search in one lane-private round takes a second look at its candidates and returns the
// four nearest as a SET (proof->others) with a bound on everything outside it: after a small move the caller
// can then settle the point among four candidates instead of searching again.
template <int TPB>
__device__ __forceinline__ int nn_search_hybrid(const GridDev &G, bool valid, const QueryCtx &c, double qx, double qy,
                                                double qz, double r2, float r2_ub, int prior, LaneRuns<TPB> &L,
                                                double *d2_out, float slack = 0.0f, SearchProof *proof = nullptr,
                                                int coop_lanes = 32, bool want_set = false) {
    const unsigned FULL = 0xffffffffu;
    static_assert((TPB & (TPB - 1)) == 0, "TPB must be a power of two");
    const GridParams &g = G.p;
    enum { kDone = 0, kScan = 1, kCoop = 2 };
    Screen r;
    r.best = r2_ub; r.second = 3.0e38f; r.bs = -1;
    const float max_reach2 = (kLaneMaxRho * g.fine) * (kLaneMaxRho * g.fine);
    float bound = r2_ub;  // f32 distance of a real candidate, when one is known
    float thr = (kLaneProbeRho * g.fine) * (kLaneProbeRho * g.fine);  // no bound yet: probe the nearest cells
    int mode = valid ? kScan : kDone;
    {
        // Bound = distance to the nearest of the previous matches of this lane and of its neighbours in the warp.
        // Under a rigid move the points of a patch slide together: a point that has moved a centimetre along the
        // surface is now nearest to what was its lane neighbour's match, and its own old match would give twice
        // the reach (eight times the cells).  The lanes are in spatial order, so the nearest old matches are
        // mostly the adjacent lanes': +-kNbWindow lanes (checking all 32 measured no faster than checking none).
        float4 t = make_float4(3.0e18f, 3.0e18f, 3.0e18f, 0.0f);
        if (valid && prior >= 0) t = __ldg(G.hi + prior);
        float nb = 3.0e38f;
#pragma unroll
        for (int o = -kNbWindow; o <= kNbWindow; ++o) {
            const int j = ((threadIdx.x & 31) + o) & 31;
            const float dx = c.qx - __shfl_sync(FULL, t.x, j), dy = c.qy - __shfl_sync(FULL, t.y, j),
                        dz = c.qz - __shfl_sync(FULL, t.z, j);
            nb = fminf(nb, fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
        }
        if (valid && nb < r2_ub) {
            bound = nb;
            thr = reach_of(g, bound, r2_ub);
            if (thr > max_reach2) { mode = kCoop; VB_STAT(4, 1); }
            else thr = widen(thr, slack, max_reach2);
        }
    }
    if (coop_lanes < 32) {
        // a warp most of whose lanes have a long reach (the first iterations of an alignment) is better served
        // by the shared walk: long reaches overlap, and listing big boxes lane by lane is what costs
        const float big2 = (kLaneBigRho * g.fine) * (kLaneBigRho * g.fine);
        const bool big = mode == kCoop || (mode == kScan && (bound >= r2_ub || thr > big2));
        if (__popc(__ballot_sync(FULL, big)) > coop_lanes && mode == kScan) mode = kCoop;
    }
    VB_STAT(0, valid);
    VB_STAT(1, bound < r2_ub);
    if ((threadIdx.x & 31) == 0) VB_STAT(11, 1);
    float done = -1.0f;  // every cell with gap2 <= done has been scanned
    int nscans = 0, nnear = 0;  // rounds this lane scanned; candidates its latest scan found below the listing threshold
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
        if (!__any_sync(FULL, mode == kScan)) break;
        int steps = 0;
        if (mode == kScan) {
            // (first round: nothing scanned yet, the box is listed by rows; a second round lists the cells it adds)
            const int nruns = done < 0.0f ? list_rows<TPB>(G, c, thr, slack, L) : list_runs<TPB>(G, c, thr, done, slack, L);
            if (nruns > kLaneMaxRuns) {
                mode = kCoop;
                VB_STAT(5, 1);
            } else {
                VB_STAT(7, nruns);
                nnear = 0;
                steps = want_set ? scan_runs<TPB, true>(G, c, nruns, bound, L, r, thr, nnear)
                                 : scan_runs<TPB, false>(G, c, nruns, bound, L, r, thr, nnear);
                done = thr;
                ++nscans;
                if (r.bs >= 0) {
                    // complete once the reach of the final best lies inside what has been scanned
                    const float need = reach_of(g, fminf(bound, r.best), r2_ub);
                    if (need <= done) mode = kDone;
                    else if (need <= max_reach2) thr = widen(need, slack, max_reach2);
                    else { mode = kCoop; VB_STAT(4, 1); }
                } else if (done < max_reach2) {
                    thr = max_reach2;  // nothing within the probe: everything this path may search
                } else {
                    mode = kCoop;
                    VB_STAT(6, 1);
                }
            }
        }
#ifdef VB_STATS
        VB_STAT(8, steps);
        {
            const int mx = __reduce_max_sync(FULL, steps);
            if ((threadIdx.x & 31) == 0) VB_STAT(9, mx);
        }
#else
        (void)steps;
#endif
    }
    if (mode == kScan) mode = kCoop;  // still open after two rounds
    const bool lane_private = mode == kDone && valid;  // every cell with gap2 <= done was scanned by this lane
    const unsigned coop = __ballot_sync(FULL, mode == kCoop);
    VB_STAT(3, mode == kCoop);
    if (coop && (threadIdx.x & 31) == 0) VB_STAT(10, 1);
    if (coop) {
        const Screen rc = coop_screen(G, coop, c, r2_ub);
        if (mode == kCoop) r = rc;
    }
    *d2_out = 0.0;
    if (proof) {
        proof->sec1 = -1.0f; proof->secK = -1.0f;
        proof->others[0] = proof->others[1] = proof->others[2] = -1;
    }
    if (!valid || r.bs < 0) return -1;
    // ---- the decision, in double
    const float bb = band(g, r.best);
    const bool unique = r.second - r.best > bb + band(g, r.second);  // r.bs is the true nearest
    // what this lane is known to have scanned: every cell within `cover`.  Points of cells never listed lie
    // beyond `done` (the gap test is deflated); after the cooperative walk only the reach of the best is known.
    const float reach = reach_of(g, fminf(bound, r.best), r2_ub);
    const float wide = widen(reach, slack, max_reach2);
    const float cover = lane_private ? fminf(done, wide) : reach;
    VB_STAT(20, want_set && lane_private && nscans == 1);
    VB_STAT(21, want_set && lane_private && nscans == 1 && nnear > kLaneMaxNear);
    VB_STAT(22, want_set && lane_private && nscans == 1 && nnear < 2);
    VB_STAT(23, want_set && valid && !(lane_private && nscans == 1));
    if (proof && want_set && lane_private && nscans == 1 && nnear >= 2 && nnear <= kLaneMaxNear) {
        // The (up to) four nearest of the candidates the scan found below its listing threshold, as a set.
        // Scanned points outside it have d32 >= min(e[4], threshold), hence true d2 >= that minus its band
        // (the threshold is `done` >= cover); unscanned ones lie beyond `cover`.
        const Top5 t = top5_of_near<TPB>(G, c, nnear, L);
        VB_STAT(16, 1);
        const float m5 = fminf(t.e[4], cover);
        const float secK = m5 - band(g, m5);
        const float ub0 = t.e[0] + band(g, t.e[0]);  // the f32-nearest candidate is truly no farther than this
        if (t.s[1] >= 0 && secK > ub0) {
            // the true nearest is in the set; contenders are the members that could be nearer than t.s[0].
            // Exact ties break to the lowest ORIGINAL index (the documented rule, as nn_exact_rescan).
            int w = 0;
            double dw = l2_exact(qx, qy, qz, G.xyz + kPtStride * (int64_t)t.s[0]);
#pragma unroll
            for (int j = 1; j < 4; j++) {
                if (t.s[j] >= 0 && t.e[j] - band(g, t.e[j]) <= ub0) {
                    const double dj = l2_exact(qx, qy, qz, G.xyz + kPtStride * (int64_t)t.s[j]);
                    if (dj < dw || (dj == dw && __ldg(G.orig + t.s[j]) < __ldg(G.orig + t.s[w]))) { w = j; dw = dj; }
                }
            }
            if (!(dw < r2)) return -1;
            *d2_out = dw;
            float o1 = 3.0e38f;  // lower bound of the true d2 of the set's other members
            int k = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (j != w) {
                    if (t.s[j] >= 0) o1 = fminf(o1, t.e[j] - band(g, t.e[j]));
                    proof->others[k++] = t.s[j];
                }
            }
            proof->secK = secK;
            proof->sec1 = fminf(secK, o1);
            VB_STAT(17, 1);
            VB_STAT(18, t.e[4] < cover);  // the set's bound is the fifth candidate, not the scanned radius
            return t.s[w];
        }
    }
    if (unique) {
        const double d = l2_exact(qx, qy, qz, G.xyz + kPtStride * (int64_t)r.bs);
        if (!(d < r2)) return -1;
        *d2_out = d;
        if (proof) {
            // scanned points other than r.bs: d32 >= r.second, hence true d2 >= r.second - band(r.second)
            const float m1 = fminf(r.second, cover);
            proof->sec1 = m1 - band(g, m1);
        }
        return r.bs;
    }
    return nn_exact_rescan(G, c, qx, qy, qz, r2, fminf(r.best + 2.0f * bb, r2_ub), d2_out);
}

// ---- warp-per-query search --------------------------------------------------------------------------------
// For few, scattered queries (the KNN sweep: 10 000 queries in a 10 M-point scene) there is no spatial
// coherence between lanes to exploit and too few queries to fill the GPU with one thread each.  Here the whole
// warp serves ONE query: the cell walk is warp-uniform, the 32 lanes take consecutive candidates of a cell
// (one coalesced 512-byte load per 32 candidates) and the per-lane best / runner-up are merged with shuffles.
// Same decisions, same exact-double fallback, same results as the other two searches.
__device__ __forceinline__ void wpq_merge(Screen &r) {
    // after this every lane holds the warp-wide best (lowest position on exact f32 ties) and runner-up
    const unsigned FULL = 0xffffffffu;
    float best = r.best;
    int bs = r.bs;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const float ob = __shfl_xor_sync(FULL, best, o);
        const int os = __shfl_xor_sync(FULL, bs, o);
        const bool take = ob < best || (ob == best && (unsigned)os < (unsigned)bs);
        best = take ? ob : best;
        bs = take ? os : bs;
    }
    // runner-up: the smallest value among every lane's runner-up and every non-winning lane's best
    float cand = (r.bs == bs) ? r.second : fminf(r.best, r.second);
#pragma unroll
    for (int o = 16; o; o >>= 1) cand = fminf(cand, __shfl_xor_sync(FULL, cand, o));
    r.best = best; r.bs = bs; r.second = cand;
}

__device__ __forceinline__ void wpq_scan_run(const float4 *__restrict__ hi, int s0, int s1, const QueryCtx &c,
                                             Screen &r) {
    const int lane = threadIdx.x & 31;
    for (int s = s0 + lane; s < s1; s += 32) {
        const float4 t = __ldg(hi + s);
        const float dx = c.qx - t.x, dy = c.qy - t.y, dz = c.qz - t.z;
        const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        const bool lt = d < r.best;
        r.second = lt ? r.best : fminf(r.second, d);
        r.bs = lt ? s : r.bs;
        r.best = fminf(r.best, d);
    }
}

// All 32 lanes call with the SAME query (c, q are warp-uniform).  Returns the same value in every lane.
__device__ __forceinline__ int nn_search_wpq(const GridDev &G, const QueryCtx &c, double qx, double qy, double qz,
                                             double r2, float r2_ub, double *d2_out) {
    const GridParams &g = G.p;
    *d2_out = 0.0;
    Screen r;  // per-lane partial state
    r.best = r2_ub; r.second = 3.0e38f; r.bs = -1;
    // home cell
    {
        const int hcx = c.gx >> 2, hcy = c.gy >> 2, hcz = c.gz >> 2;
        const CoarseCell cc = G.coarse[((int64_t)hcz * g.cdim[1] + hcy) * g.cdim[0] + hcx];
        const int hbit = (c.gx & 3) + 4 * (c.gy & 3) + 16 * (c.gz & 3);
        if ((cc.mask >> hbit) & 1ull) {
            const int rank = __popcll(cc.mask & ((1ull << hbit) - 1ull));
            wpq_scan_run(G.hi, __ldg(G.fstart + cc.base + rank), __ldg(G.fstart + cc.base + rank + 1), c, r);
        }
    }
    Screen m = r;
    wpq_merge(m);
    float thr = reach_of(g, m.best, r2_ub);
    auto visit = [&](int s0, int s1, float gap2) {
        if (gap2 > thr) return;  // warp-uniform: thr derives from the merged best
        wpq_scan_run(G.hi, s0, s1, c, r);
        if (s1 - s0 > 0) {
            m = r;
            wpq_merge(m);
            thr = reach_of(g, m.best, r2_ub);
        }
    };
    walk_cells(G, c, thr, true, visit);
    m = r;
    wpq_merge(m);
    if (m.bs < 0) return -1;
    const float bb = band(g, m.best);
    int out;
    double d2 = 0.0;
    if (m.second - m.best > bb + band(g, m.second)) {
        const double d = l2_exact(qx, qy, qz, G.xyz + kPtStride * (int64_t)m.bs);
        out = d < r2 ? m.bs : -1;
        d2 = d < r2 ? d : 0.0;
    } else {
        out = nn_exact_rescan(G, c, qx, qy, qz, r2, fminf(m.best + 2.0f * bb, r2_ub), &d2);  // uniform: every lane repeats it
    }
    *d2_out = d2;
    return out;
}

// ---- warp-per-query search, breadth first ---------------------------------------------------------------
// nn_search_wpq above walks the cells one after the other: coarse cell -> fine-cell range -> candidates, each a
// dependent load, 30-50 DRAM round trips per query when nothing is cached (the KNN sweep: 41 us for 10 000
// queries whatever the scene size).  Here the same search is laid out in a handful of round trips:
//   0. home cell (coarse cell, range, candidates): a first bound, hence the reach `thr` of everything below;
//   1. the <= 27 coarse cells of the reach box, one per lane, in ONE round trip; each lane prunes its cell's
//      occupied fine cells against `thr` (same deflated gap test as walk_cells);
//   2. the surviving cells' ranges, 32 per round trip, into a per-warp list in shared memory;
//   3. their candidates as ONE flat list the lanes stride over, 128 loads in flight per round trip.
// `thr` is not tightened while scanning, so a few more candidates are evaluated than the sequential walk would
// (never fewer: every cell within reach_of(final best) <= thr is scanned) — same Screen semantics, same decision,
// same results bit for bit.  More than kWpqMaxRuns cells (cannot happen below ~2^8 occupied cells within the
// radius) falls back to the sequential walk.
constexpr int kWpqMaxRuns = 256;
struct WpqScratch {
    int s0[kWpqMaxRuns];
    int len[kWpqMaxRuns];
};

__device__ __forceinline__ int nn_search_wpq_bfs(const GridDev &G, const QueryCtx &c, double qx, double qy, double qz,
                                                 double r2, float r2_ub, WpqScratch &ws, double *d2_out) {
    const unsigned FULL = 0xffffffffu;
    const GridParams &g = G.p;
    const int lane = threadIdx.x & 31;
    *d2_out = 0.0;
    Screen r;  // per-lane partial state
    r.best = r2_ub; r.second = 3.0e38f; r.bs = -1;
    const int hcx = c.gx >> 2, hcy = c.gy >> 2, hcz = c.gz >> 2;
    const int hbit = (c.gx & 3) + 4 * (c.gy & 3) + 16 * (c.gz & 3);
    {
        const CoarseCell cc = G.coarse[((int64_t)hcz * g.cdim[1] + hcy) * g.cdim[0] + hcx];
        if ((cc.mask >> hbit) & 1ull) {
            const int rank = __popcll(cc.mask & ((1ull << hbit) - 1ull));
            wpq_scan_run(G.hi, __ldg(G.fstart + cc.base + rank), __ldg(G.fstart + cc.base + rank + 1), c, r);
        }
    }
    Screen m = r;
    wpq_merge(m);
    const float thr = reach_of(g, m.best, r2_ub);  // warp-uniform; the whole radius when the home cell is empty
    // the reach box in fine cells (as walk_cells)
    const float fine = g.fine, fine2 = fine * fine;
    const float rho = sqrtf(thr) / fine * 1.00001f + 1e-6f;
    const int fx0 = max(c.gx - (int)ceilf(fmaxf(rho - c.fx, 0.0f)), 0),
              fx1 = min(c.gx + (int)ceilf(fmaxf(rho - (1.0f - c.fx), 0.0f)), g.fdim[0] - 1);
    const int fy0 = max(c.gy - (int)ceilf(fmaxf(rho - c.fy, 0.0f)), 0),
              fy1 = min(c.gy + (int)ceilf(fmaxf(rho - (1.0f - c.fy), 0.0f)), g.fdim[1] - 1);
    const int fz0 = max(c.gz - (int)ceilf(fmaxf(rho - c.fz, 0.0f)), 0),
              fz1 = min(c.gz + (int)ceilf(fmaxf(rho - (1.0f - c.fz), 0.0f)), g.fdim[2] - 1);
    const int ncx = (fx1 >> 2) - (fx0 >> 2) + 1, ncy = (fy1 >> 2) - (fy0 >> 2) + 1, ncz = (fz1 >> 2) - (fz0 >> 2) + 1;
    if (ncx * ncy * ncz > 32) return nn_search_wpq(G, c, qx, qy, qz, r2, r2_ub, d2_out);  // reach <= coarse cell: <= 27
    // 1. one coarse cell per lane
    unsigned long long sel = 0ull, msk = 0ull;
    int base = 0;
    if (lane < ncx * ncy * ncz) {
        const int cx = (fx0 >> 2) + lane % ncx, cy = (fy0 >> 2) + (lane / ncx) % ncy, cz = (fz0 >> 2) + lane / (ncx * ncy);
        const CoarseCell cc = G.coarse[((int64_t)cz * g.cdim[1] + cy) * g.cdim[0] + cx];
        msk = cc.mask;
        base = cc.base;
        unsigned long long cand = msk & range_mask(max(fx0 - 4 * cx, 0), min(fx1 - 4 * cx, 3), max(fy0 - 4 * cy, 0),
                                                   min(fy1 - 4 * cy, 3), max(fz0 - 4 * cz, 0), min(fz1 - 4 * cz, 3));
        if (cx == hcx && cy == hcy && cz == hcz) cand &= ~(1ull << hbit);
        const float rx = (float)(c.gx - 4 * cx) + c.fx, ry = (float)(c.gy - 4 * cy) + c.fy, rz = (float)(c.gz - 4 * cz) + c.fz;
        while (cand) {
            const int b = __ffsll((long long)cand) - 1;
            cand &= cand - 1ull;
            const float fx = small_int_to_float(b & 3), fy = small_int_to_float((b >> 2) & 3), fz = small_int_to_float(b >> 4);
            const float ex = fmaxf(fmaxf(fx - rx, rx - fx - 1.0f), 0.0f);
            const float ey = fmaxf(fmaxf(fy - ry, ry - fy - 1.0f), 0.0f);
            const float ez = fmaxf(fmaxf(fz - rz, rz - fz - 1.0f), 0.0f);
            const float gap2 = (ex * ex + ey * ey + ez * ez) * fine2 * 0.998f - 1e-12f * fine2;  // deflated lower bound
            if (gap2 <= thr) sel |= 1ull << b;
        }
    }
    // 2. the surviving cells as one list: positions in fstart first, then (start, length) in one round trip per 32
    const int mine = __popcll(sel);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    const int nruns = __shfl_sync(FULL, incl, 31);
    if (nruns > kWpqMaxRuns) return nn_search_wpq(G, c, qx, qy, qz, r2, r2_ub, d2_out);
    {
        int o = incl - mine;
        unsigned long long t = sel;
        while (t) {
            const int b = __ffsll((long long)t) - 1;
            t &= t - 1ull;
            ws.s0[o++] = base + __popcll(msk & ((1ull << b) - 1ull));
        }
    }
    __syncwarp();
    for (int i = lane; i < nruns; i += 32) {
        const int f = ws.s0[i];
        const int a = __ldg(G.fstart + f), e = __ldg(G.fstart + f + 1);
        ws.s0[i] = a;
        ws.len[i] = e - a;
    }
    __syncwarp();
    // 3. candidates: 32 runs at a time form a flat list (occupied cells are never empty, so the runs' ends are
    // strictly increasing and a ballot finds each index's run); four windows of 32 loads are issued together
    const unsigned le_mask = 0xffffffffu >> (31 - lane);
    for (int g0 = 0; g0 < nruns; g0 += 32) {
        const bool have = g0 + lane < nruns;
        const int rs0 = have ? ws.s0[g0 + lane] : 0, rlen = have ? ws.len[g0 + lane] : 0;
        int rend = rlen;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, rend, o);
            if (lane >= o) rend += v;
        }
        const int ncand = __shfl_sync(FULL, rend, 31);
        const int rstart = rend - rlen;
        const int my_end = rlen > 0 ? rend : -1;
        int kbase = 0;
        for (int t0 = 0; t0 < ncand; t0 += 128) {
            float4 cv[4];
            int cs[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int w0 = t0 + 32 * u;
                const unsigned rel = (unsigned)(my_end - w0);
                const unsigned endm = __reduce_or_sync(FULL, rel < 32u ? 1u << rel : 0u);
                const int k = min(kbase + __popc(endm & le_mask), 31);
                kbase += __popc(endm);
                const int t = w0 + lane;
                const int s = __shfl_sync(FULL, rs0, k) + (t - __shfl_sync(FULL, rstart, k));
                cs[u] = t < ncand ? s : -1;
                if (t < ncand) cv[u] = __ldg(G.hi + s);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (cs[u] >= 0) {
                    const float dx = c.qx - cv[u].x, dy = c.qy - cv[u].y, dz = c.qz - cv[u].z;
                    const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    const bool lt = d < r.best;
                    r.second = lt ? r.best : fminf(r.second, d);
                    r.bs = lt ? cs[u] : r.bs;
                    r.best = fminf(r.best, d);
                }
            }
        }
    }
    m = r;
    wpq_merge(m);
    if (m.bs < 0) return -1;
    const float bb = band(g, m.best);
    int out;
    double d2 = 0.0;
    if (m.second - m.best > bb + band(g, m.second)) {
        const double d = l2_exact(qx, qy, qz, G.xyz + kPtStride * (int64_t)m.bs);
        out = d < r2 ? m.bs : -1;
        d2 = d < r2 ? d : 0.0;
    } else {
        out = nn_exact_rescan(G, c, qx, qy, qz, r2, fminf(m.best + 2.0f * bb, r2_ub), &d2);  // uniform: every lane repeats it
    }
    *d2_out = d2;
    return out;
}

#endif  // __CUDACC__


}  // namespace vb
