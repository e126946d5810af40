// Rank fusion kernels: HybridRetriever.reciprocal_rank_fusion (retrievers.py:256-274)
// and HybridRetriever.fusion (retrievers.py:239-253), one CTA per query.
//
// Both reference functions key on the chunk TEXT (node.get_content()), not the
// node id; canon[id] is the smallest document index carrying the same text, so
// integer keys reproduce the dict semantics exactly:
//   * rrf_map[text] += 1/(rank+K) in list order (a first, then b), float64;
//   * text_to_node[text] = item  -> the LAST occurrence supplies the node;
//   * sorted(..., reverse=True) is stable -> ties keep first-insertion order.
// Lists are at most a few hundred entries (f_topk 288/192), so the O(n^2)
// dedup/rank below is a handful of shared-memory sweeps; the work is latency
// bound and excluded from the roofline fraction.
#include "ezr_common.cuh"
#include "../../include/easyrag_b200.h"

namespace ezr {

constexpr int kFuseThreads = 128;
constexpr int kFuseMaxIn = 1024;   // entries per input list

template <bool RRF>
__global__ void __launch_bounds__(kFuseThreads)
fuse_kernel(const int32_t* __restrict__ ids_a, const double* __restrict__ sc_a, const int32_t* __restrict__ cnt_a,
            const int32_t* __restrict__ ids_b, const double* __restrict__ sc_b, const int32_t* __restrict__ cnt_b,
            int stride_in, const int32_t* __restrict__ canon, int canon_base, int K, int k_out,
            int32_t* __restrict__ out_ids, double* __restrict__ out_scores, int32_t* __restrict__ out_counts) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int q = blockIdx.x;
    const int ca = min(max(cnt_a[q], 0), stride_in);
    const int cb = min(max(cnt_b[q], 0), stride_in);
    const int n = ca + cb;
    // carve: key[n] id[n] rep[n] lead[n] | score[n]
    const int cap = 2 * stride_in;
    double* s_sc = reinterpret_cast<double*>(smem_raw);
    int* s_key = reinterpret_cast<int*>(smem_raw + (size_t)cap * 8);
    int* s_id = s_key + cap;
    int* s_rep = s_id + cap;
    int* s_lead = s_rep + cap;
    __shared__ int s_nlead;
    if (threadIdx.x == 0) s_nlead = 0;

    for (int e = threadIdx.x; e < n; e += kFuseThreads) {
        const bool in_a = e < ca;
        const int j = in_a ? e : e - ca;
        const int id = in_a ? ids_a[(int64_t)q * stride_in + j] : ids_b[(int64_t)q * stride_in + j];
        s_id[e] = id;
        s_key[e] = (canon && id >= 0) ? canon[id - canon_base] : id;
        if (!RRF) s_sc[e] = in_a ? sc_a[(int64_t)q * stride_in + j] : sc_b[(int64_t)q * stride_in + j];
    }
    __syncthreads();

    for (int e = threadIdx.x; e < n; e += kFuseThreads) {
        const int key = s_key[e];
        bool first = true;
        for (int j = 0; j < e; ++j)
            if (s_key[j] == key) { first = false; break; }
        s_lead[e] = first ? 1 : 0;
        if (first) {
            if (RRF) {
                double acc = 0.0;
                int last = e;
                for (int j = e; j < n; ++j) {
                    if (s_key[j] == key) {
                        const int rank = (j < ca ? j : j - ca) + 1;
                        acc = __dadd_rn(acc, __ddiv_rn(1.0, (double)(rank + K)));
                        last = j;
                    }
                }
                s_sc[e] = acc;
                s_rep[e] = s_id[last];     // text_to_node: last writer wins
            } else {
                s_rep[e] = s_id[e];        // fusion(): first occurrence is kept
            }
            atomicAdd(&s_nlead, 1);
        }
    }
    __syncthreads();

    for (int e = threadIdx.x; e < n; e += kFuseThreads) {
        if (!s_lead[e]) continue;
        const double sc = s_sc[e];
        int pos = 0;
        for (int j = 0; j < n; ++j) {
            if (!s_lead[j]) continue;
            const double sj = s_sc[j];
            pos += (sj > sc || (sj == sc && j < e)) ? 1 : 0;   // stable descending
        }
        if (pos < k_out) {
            out_ids[(int64_t)q * k_out + pos] = s_rep[e];
            out_scores[(int64_t)q * k_out + pos] = sc;
        }
    }
    const int nl = min(s_nlead, k_out);
    for (int i = nl + threadIdx.x; i < k_out; i += kFuseThreads) {
        out_ids[(int64_t)q * k_out + i] = -1;
        out_scores[(int64_t)q * k_out + i] = -INFINITY;
    }
    if (threadIdx.x == 0) out_counts[q] = nl;
}

template <bool RRF>
static int fuse_launch(const int32_t* ids_a, const double* sc_a, const int32_t* cnt_a, const int32_t* ids_b,
                       const double* sc_b, const int32_t* cnt_b, int n_queries, int stride_in,
                       const int32_t* canon, int canon_base, int K, int k_out, int32_t* out_ids,
                       double* out_scores, int32_t* out_counts, cudaStream_t st) {
    EZR_CHECK_ARG(stride_in >= 1 && stride_in <= kFuseMaxIn, "fusion: stride_in=%d out of [1,%d]", stride_in,
                  kFuseMaxIn);
    EZR_CHECK_ARG(k_out >= 1, "fusion: k_out must be >= 1");
    if (n_queries == 0) return EZR_OK;
    const size_t smem = (size_t)2 * stride_in * (8 + 4 * 4);
    static bool attr_done[2] = {false, false};
    if (!attr_done[RRF ? 1 : 0]) {
        EZR_CUDA(cudaFuncSetAttribute(fuse_kernel<RRF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      2 * kFuseMaxIn * 24));
        attr_done[RRF ? 1 : 0] = true;
    }
    ProfScope prof(EZR_PROF_FUSE, st);
    fuse_kernel<RRF><<<n_queries, kFuseThreads, smem, st>>>(ids_a, sc_a, cnt_a, ids_b, sc_b, cnt_b, stride_in, canon,
                                                            canon_base, K, k_out, out_ids, out_scores, out_counts);
    EZR_LAUNCH_CHECK();
    return EZR_OK;
}

}  // namespace ezr

using namespace ezr;

extern "C" {

int ezr_rrf_fuse(const int32_t* ids_a, const int32_t* cnt_a, const int32_t* ids_b, const int32_t* cnt_b,
                 int32_t n_queries, int32_t stride_in, const int32_t* canon, int32_t canon_base, int32_t K,
                 int32_t k_out, int32_t* out_ids, double* out_scores, int32_t* out_counts, void* stream) {
    EZR_CHECK_ARG(K >= 0, "rrf: K must be >= 0");
    return fuse_launch<true>(ids_a, nullptr, cnt_a, ids_b, nullptr, cnt_b, n_queries, stride_in, canon, canon_base,
                             K, k_out, out_ids, out_scores, out_counts, (cudaStream_t)stream);
}

int ezr_fusion_simple(const int32_t* ids_a, const double* scores_a, const int32_t* cnt_a, const int32_t* ids_b,
                      const double* scores_b, const int32_t* cnt_b, int32_t n_queries, int32_t stride_in,
                      const int32_t* canon, int32_t canon_base, int32_t k_out, int32_t* out_ids,
                      double* out_scores, int32_t* out_counts, void* stream) {
    return fuse_launch<false>(ids_a, scores_a, cnt_a, ids_b, scores_b, cnt_b, n_queries, stride_in, canon,
                              canon_base, 0, k_out, out_ids, out_scores, out_counts, (cudaStream_t)stream);
}

}  // extern "C"
