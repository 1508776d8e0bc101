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
#endif

}  // namespace vb
