// Batched Hamming DBSCAN on bit matrices (order-independent restatement of
// sklearn.cluster.DBSCAN(metric="hamming") as used at identify_visibility_windows.py:114 and
// cotracker_matching.py:809; semantics in SURVEY.md Appendix A.6).
#pragma once
#include "common.cuh"

namespace s2d {

struct DbProblem {
    const uint32_t* bits;   // row 0 of the problem
    const uint8_t* valid;   // optional per-row flag; rows with 0 are all-zero rows
    int32_t* core;          // work [N]
    int32_t* parent;        // work [N]
    int32_t* aux;           // work [N]
    int32_t* labels;        // out  [N]
    int32_t* nclusters;     // out  (optional)
    const int32_t* gram;    // optional: G = X X^T of the rows (row 0 of the problem at gram[0], pitch gstride); distances
    int32_t gstride, pad;   //           then come from G instead of XOR + popc over the bit rows
    int32_t stride;         // words per row
    int32_t w0, nw;         // word window compared
    int32_t N;
    int32_t kmax;           // neighbours <=> hamming <= kmax
    int32_t min_samples;
};

// largest k in [0, D] with double(k)/double(D) <= eps, -1 if none (python: float(k)/float(D))
__device__ __forceinline__ int hamming_kmax(int D, double eps) {
    if (D <= 0) return 0;
    int k = (int)floor(eps * (double)D) + 2;
    if (k > D) k = D;
    while (k >= 0 && !((double)k / (double)D <= eps)) --k;
    return k;
}

// launches the four phases for `nproblems` problems whose descriptors live in device memory;
// max_N bounds every problem's N.
// `wl`: int32 scratch of db_worklist_ints(nproblems) elements (the passes' work list, built on the device).
inline int64_t db_worklist_ints(int nproblems) { return 2 * ((int64_t)nproblems + 1) + 2; }
int launch_dbscan(const DbProblem* problems, int nproblems, int max_N, int max_nw, int32_t* wl, cudaStream_t st);

}  // namespace s2d
