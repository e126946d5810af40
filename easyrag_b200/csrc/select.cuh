// Block-level top-k selection for arbitrary k <= 1024 (threshold + candidate
// buffer + shared-memory bitonic sort).  Used where k exceeds what WarpTopK
// (k<=32) covers: the drop-in retrievers run with k = 288/256/192
// (easyrag.yaml:8-11).  The result is the unique top-k under the canonical
// order, independent of the order candidates were pushed in.
#pragma once
#include "ezr_common.cuh"

namespace ezr {

constexpr int kSelMaxK = 1024;
constexpr int kSelReserve = 1024;                       // max pushes between two flush points
constexpr int kSelCap = kSelMaxK + 2 * kSelReserve;     // 3072
constexpr int kSelPow2 = 4096;

template <typename S>
struct SelSmem {
    S* ks;
    int* kid;
    int* cnt;      // [0] count
    S* thr_s;      // [0]
    int* thr_id;   // [0]
};

template <typename S>
__host__ __device__ inline size_t sel_smem_bytes() {
    return (size_t)kSelPow2 * (sizeof(S) + sizeof(int)) + 64;
}

template <typename S>
__device__ __forceinline__ SelSmem<S> sel_carve(unsigned char* base) {
    SelSmem<S> m;
    m.ks = reinterpret_cast<S*>(base);
    m.kid = reinterpret_cast<int*>(base + (size_t)kSelPow2 * sizeof(S));
    unsigned char* tail = base + (size_t)kSelPow2 * (sizeof(S) + sizeof(int));
    m.thr_s = reinterpret_cast<S*>(tail);            // 8B aligned: kSelPow2*(S+4) is a multiple of 8
    m.cnt = reinterpret_cast<int*>(tail + 16);
    m.thr_id = reinterpret_cast<int*>(tail + 20);
    return m;
}

template <typename S>
__device__ __forceinline__ void sel_init(const SelSmem<S>& m) {
    if (threadIdx.x == 0) {
        *m.cnt = 0;
        *m.thr_s = ScoreTraits<S>::lowest();
        *m.thr_id = -1;
    }
    __syncthreads();
}

template <typename S>
__device__ __forceinline__ void sel_push(const SelSmem<S>& m, S s, int id) {
    if (better<S>(s, id, *m.thr_s, *m.thr_id)) {
        const int p = atomicAdd(m.cnt, 1);
        m.ks[p] = s;
        m.kid[p] = id;
    }
}

template <typename S>
__device__ void bitonic_sort_desc(S* ks, int* kid, int P) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const S a = ks[lo], b = ks[hi];
                const int ia = kid[lo], ib = kid[hi];
                const bool sw = desc ? better<S>(b, ib, a, ia) : better<S>(a, ia, b, ib);
                if (sw) {
                    ks[lo] = b; ks[hi] = a;
                    kid[lo] = ib; kid[hi] = ia;
                }
            }
            __syncthreads();
        }
    }
}

// Sort what is buffered, keep the best k, raise the threshold.  All threads call.
template <typename S>
__device__ void sel_compact(const SelSmem<S>& m, int k) {
    __syncthreads();
    const int n = *m.cnt;
    int P = 2;
    while (P < n) P <<= 1;
    for (int i = n + threadIdx.x; i < P; i += blockDim.x) {
        m.ks[i] = ScoreTraits<S>::lowest();
        m.kid[i] = -1;
    }
    __syncthreads();
    bitonic_sort_desc<S>(m.ks, m.kid, P);
    if (threadIdx.x == 0) {
        const int keep = n < k ? n : k;
        *m.cnt = keep;
        if (n >= k) {
            *m.thr_s = m.ks[k - 1];
            *m.thr_id = m.kid[k - 1];
        }
    }
    __syncthreads();
}

// Call at block-uniform points, at most kSelReserve pushes apart.
template <typename S>
__device__ __forceinline__ void sel_maybe_flush(const SelSmem<S>& m, int k) {
    __syncthreads();
    const int n = *m.cnt;
    __syncthreads();          // nobody may push again before every thread has read the count
    if (n > kSelCap - kSelReserve) sel_compact<S>(m, k);
}

}  // namespace ezr
