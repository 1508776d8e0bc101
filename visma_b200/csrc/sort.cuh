// sort.cuh — single-thread in-place sort used to put each grid cell / bucket into a deterministic order
// after the atomic counting-sort scatter.  Cells hold tens to hundreds of keys.
#pragma once

namespace vb {

#ifdef __CUDACC__
template <typename K>
__device__ inline void cell_sort(K *a, int n) {
    if (n < 2) return;
    if (n <= 16) {
        for (int i = 1; i < n; i++) {
            K v = a[i];
            int j = i - 1;
            while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; j--; }
            a[j + 1] = v;
        }
        return;
    }
    for (int start = n / 2 - 1; start >= 0; start--) {
        int root = start;
        K v = a[root];
        for (;;) {
            int child = 2 * root + 1;
            if (child >= n) break;
            if (child + 1 < n && a[child + 1] > a[child]) child++;
            if (a[child] <= v) break;
            a[root] = a[child];
            root = child;
        }
        a[root] = v;
    }
    for (int end = n - 1; end > 0; end--) {
        K v = a[end];
        a[end] = a[0];
        int root = 0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && a[child + 1] > a[child]) child++;
            if (a[child] <= v) break;
            a[root] = a[child];
            root = child;
        }
        a[root] = v;
    }
}

// Warp-cooperative sort of one cell: every lane keeps up to MAXPL keys in registers, ranks them against all n
// keys (warp-uniform broadcast loads: n * MAXPL compares per lane) and scatters them to their rank.  Keys are
// unique (they embed the point index), so ranks are a permutation.  Cells beyond 32*MAXPL keys fall back to
// the single-thread heapsort on lane 0.  All 32 lanes must call.
template <typename K, int MAXPL>
__device__ inline void warp_cell_sort(K *a, int n) {
    const int lane = threadIdx.x & 31;
    if (n < 2) return;
    if (n > 32 * MAXPL) {
        if (lane == 0) cell_sort(a, n);
        __syncwarp();
        return;
    }
    K mine[MAXPL];
    int rank[MAXPL];
#pragma unroll
    for (int k = 0; k < MAXPL; k++) {
        const int i = lane + 32 * k;
        mine[k] = i < n ? a[i] : K(0);
        rank[k] = 0;
    }
    for (int j = 0; j < n; j++) {
        const K v = a[j];
#pragma unroll
        for (int k = 0; k < MAXPL; k++) rank[k] += (v < mine[k]) ? 1 : 0;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < MAXPL; k++) {
        const int i = lane + 32 * k;
        if (i < n) a[rank[k]] = mine[k];
    }
    __syncwarp();
}
#endif

}  // namespace vb
