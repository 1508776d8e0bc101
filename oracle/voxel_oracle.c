/* oracle/voxel_oracle.c — CPU restatement of open3d::VoxelDownSample.  TEST INFRASTRUCTURE ONLY.
 * Follows O3D/src/Core/Geometry/DownSample.cpp:179-220 and AccumulatedPoint (:38-87); pinned by the
 * unmodified reference (oracle/_ref) in tests/test_oracle_vs_ref.py.
 */
#include "oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int32_t v[3]; int32_t idx; } vkey;

static int vkey_cmp(const void *a, const void *b) {
    const vkey *x = (const vkey *)a, *y = (const vkey *)b;
    for (int k = 2; k >= 0; k--)
        if (x->v[k] != y->v[k]) return x->v[k] < y->v[k] ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx); /* keep input order inside a voxel */
}

int64_t vo_voxel_downsample(const double *xyz, const double *nrm, int64_t n, double voxel,
                            double *out_xyz, double *out_nrm) {
    if (voxel <= 0.0) return -1;                                      /* :183-186 */
    if (n <= 0) return 0;
    double lo[3], hi[3];
    for (int a = 0; a < 3; a++) { lo[a] = xyz[a]; hi[a] = xyz[a]; }
    for (int64_t i = 0; i < n; i++)
        for (int a = 0; a < 3; a++) {
            double v = xyz[3 * i + a];
            if (v < lo[a]) lo[a] = v;
            if (v > hi[a]) hi[a] = v;
        }
    double mn[3], ext = 0.0;
    for (int a = 0; a < 3; a++) {
        mn[a] = lo[a] - voxel * 0.5;                                  /* :189 */
        double mx = hi[a] + voxel * 0.5;                              /* :190 */
        if (mx - mn[a] > ext) ext = mx - mn[a];
    }
    if (voxel * (double)INT_MAX < ext) return -1;                     /* :191-195 */
    vkey *keys = (vkey *)malloc(sizeof(vkey) * (size_t)n);
    for (int64_t i = 0; i < n; i++) {
        for (int a = 0; a < 3; a++)
            keys[i].v[a] = (int32_t)floor((xyz[3 * i + a] - mn[a]) / voxel); /* :201-203 */
        keys[i].idx = (int32_t)i;
    }
    qsort(keys, (size_t)n, sizeof(vkey), vkey_cmp);
    /* voxels are emitted in order of FIRST APPEARANCE in the input (the reference's order is the
     * unordered_map's, i.e. unspecified; callers compare as sets).  rank[i] = output slot of the voxel whose
     * first point is i. */
    int64_t *rank = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    memset(rank, 0, sizeof(int64_t) * (size_t)n);
    for (int64_t s = 0; s < n; s++)
        if (s == 0 || keys[s].v[0] != keys[s - 1].v[0] || keys[s].v[1] != keys[s - 1].v[1] ||
            keys[s].v[2] != keys[s - 1].v[2])
            rank[keys[s].idx] = 1;
    {
        int64_t run = 0;
        for (int64_t i = 0; i < n; i++) { int64_t f = rank[i]; rank[i] = run; run += f; }
    }
    int64_t nv = 0;
    for (int64_t s = 0; s < n;) {
        int64_t e = s;
        double p[3] = {0, 0, 0}, q[3] = {0, 0, 0};
        while (e < n && keys[e].v[0] == keys[s].v[0] && keys[e].v[1] == keys[s].v[1] &&
               keys[e].v[2] == keys[s].v[2]) {
            int64_t i = keys[e].idx;
            for (int a = 0; a < 3; a++) p[a] += xyz[3 * i + a];       /* AddPoint :50 */
            if (nrm && !isnan(nrm[3 * i]) && !isnan(nrm[3 * i + 1]) && !isnan(nrm[3 * i + 2]))
                for (int a = 0; a < 3; a++) q[a] += nrm[3 * i + a];   /* :51-57 */
            e++;
        }
        double cnt = (double)(e - s);
        int64_t o = rank[keys[s].idx];
        for (int a = 0; a < 3; a++) out_xyz[3 * o + a] = p[a] / cnt; /* GetAveragePoint :66 */
        if (nrm && out_nrm) {
            double l = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);  /* normalized() :71 */
            for (int a = 0; a < 3; a++) out_nrm[3 * o + a] = l > 0.0 ? q[a] / l : q[a];
        }
        nv++;
        s = e;
    }
    free(keys);
    free(rank);
    return nv;
}
