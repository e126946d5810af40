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

constexpr int kFuseMaxLists = 8;   // rank lists per call (the pipeline passes two: pipeline.py:362,408)
constexpr int kFuseMaxTotal = 2048; // entries over all lists

struct FuseLists {
    const int32_t* ids[kFuseMaxLists];
    const double* sc[kFuseMaxLists];
    const int32_t* cnt[kFuseMaxLists];
    int n;
};

template <bool RRF>
__global__ void __launch_bounds__(kFuseThreads)
fuse_kernel(const FuseLists lists, int stride_in, const int32_t* __restrict__ canon, int canon_base, int K, int k_out,
            int32_t* __restrict__ out_ids, double* __restrict__ out_scores, int32_t* __restrict__ out_counts) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int q = blockIdx.x;
    // off[l] = entries of the lists before list l (list order = insertion order of the reference's loops)
    int off[kFuseMaxLists + 1];
    off[0] = 0;
#pragma unroll
    for (int l = 0; l < kFuseMaxLists; ++l)
        off[l + 1] = off[l] + (l < lists.n ? min(max(lists.cnt[l][q], 0), stride_in) : 0);
    const int n = off[kFuseMaxLists];
    // carve: score[cap] | key[cap] id[cap] rep[cap] lead[cap] rank[cap]
    const int cap = lists.n * stride_in;
    double* s_sc = reinterpret_cast<double*>(smem_raw);
    int* s_key = reinterpret_cast<int*>(smem_raw + (size_t)cap * 8);
    int* s_id = s_key + cap;
    int* s_rep = s_id + cap;
    int* s_lead = s_rep + cap;
    int* s_rank = s_lead + cap;
    __shared__ int s_nlead;
    if (threadIdx.x == 0) s_nlead = 0;

    for (int e = threadIdx.x; e < n; e += kFuseThreads) {
        int l = 0;
#pragma unroll
        for (int t = 1; t < kFuseMaxLists; ++t) l += (e >= off[t]) ? 1 : 0;      // off is non-decreasing
        const int j = e - off[l];
        const int id = lists.ids[l][(int64_t)q * stride_in + j];
        s_id[e] = id;
        s_key[e] = (canon && id >= 0) ? canon[id - canon_base] : id;
        s_rank[e] = j + 1;                                                       // enumerate(rank_list, 1)
        if (!RRF) s_sc[e] = lists.sc[l][(int64_t)q * stride_in + j];
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
                        acc = __dadd_rn(acc, __ddiv_rn(1.0, (double)(s_rank[j] + K)));
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
static int fuse_launch(const FuseLists& lists, int n_queries, int stride_in, const int32_t* canon, int canon_base, int K,
                       int k_out, int32_t* out_ids, double* out_scores, int32_t* out_counts, cudaStream_t st) {
    EZR_CHECK_ARG(lists.n >= 1 && lists.n <= kFuseMaxLists, "fusion: %d lists (1..%d supported)", lists.n, kFuseMaxLists);
    EZR_CHECK_ARG(stride_in >= 1 && stride_in <= kFuseMaxIn, "fusion: stride_in=%d out of [1,%d]", stride_in,
                  kFuseMaxIn);
    EZR_CHECK_ARG(lists.n * stride_in <= kFuseMaxTotal, "fusion: %d lists x %d entries > %d", lists.n, stride_in,
                  kFuseMaxTotal);
    EZR_CHECK_ARG(k_out >= 1, "fusion: k_out must be >= 1");
    for (int l = 0; l < lists.n; ++l)
        EZR_CHECK_ARG(lists.ids[l] && lists.cnt[l] && (RRF || lists.sc[l]), "fusion: list %d has a NULL array", l);
    if (n_queries == 0) return EZR_OK;
    const size_t smem = (size_t)lists.n * stride_in * (8 + 5 * 4);
    static bool attr_done[2] = {false, false};
    if (!attr_done[RRF ? 1 : 0]) {
        EZR_CUDA(cudaFuncSetAttribute(fuse_kernel<RRF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kFuseMaxTotal * 28));
        attr_done[RRF ? 1 : 0] = true;
    }
    ProfScope prof(EZR_PROF_FUSE, st);
    fuse_kernel<RRF><<<n_queries, kFuseThreads, smem, st>>>(lists, stride_in, canon, canon_base, K, k_out, out_ids,
                                                            out_scores, out_counts);
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
    FuseLists l = {};
    l.n = 2;
    l.ids[0] = ids_a; l.cnt[0] = cnt_a;
    l.ids[1] = ids_b; l.cnt[1] = cnt_b;
    return fuse_launch<true>(l, n_queries, stride_in, canon, canon_base, K, k_out, out_ids, out_scores, out_counts,
                             (cudaStream_t)stream);
}

int ezr_fusion_simple(const int32_t* ids_a, const double* scores_a, const int32_t* cnt_a, const int32_t* ids_b,
                      const double* scores_b, const int32_t* cnt_b, int32_t n_queries, int32_t stride_in,
                      const int32_t* canon, int32_t canon_base, int32_t k_out, int32_t* out_ids,
                      double* out_scores, int32_t* out_counts, void* stream) {
    FuseLists l = {};
    l.n = 2;
    l.ids[0] = ids_a; l.sc[0] = scores_a; l.cnt[0] = cnt_a;
    l.ids[1] = ids_b; l.sc[1] = scores_b; l.cnt[1] = cnt_b;
    return fuse_launch<false>(l, n_queries, stride_in, canon, canon_base, 0, k_out, out_ids, out_scores, out_counts,
                              (cudaStream_t)stream);
}

int ezr_fuse_lists(int32_t rrf, int32_t n_lists, const int32_t* const* ids_host, const double* const* scores_host,
                   const int32_t* const* cnt_host, int32_t n_queries, int32_t stride_in, const int32_t* canon,
                   int32_t canon_base, int32_t K, int32_t k_out, int32_t* out_ids, double* out_scores,
                   int32_t* out_counts, void* stream) {
    EZR_CHECK_ARG(n_lists >= 1 && n_lists <= kFuseMaxLists, "fuse_lists: %d lists (1..%d supported)", n_lists,
                  kFuseMaxLists);
    EZR_CHECK_ARG(ids_host && cnt_host && (rrf || scores_host), "fuse_lists: NULL pointer table");
    EZR_CHECK_ARG(!rrf || K >= 0, "rrf: K must be >= 0");
    FuseLists l = {};
    l.n = n_lists;
    for (int i = 0; i < n_lists; ++i) {
        l.ids[i] = ids_host[i];
        l.cnt[i] = cnt_host[i];
        l.sc[i] = scores_host ? scores_host[i] : nullptr;
    }
    return rrf ? fuse_launch<true>(l, n_queries, stride_in, canon, canon_base, K, k_out, out_ids, out_scores, out_counts,
                                   (cudaStream_t)stream)
               : fuse_launch<false>(l, n_queries, stride_in, canon, canon_base, 0, k_out, out_ids, out_scores,
                                    out_counts, (cudaStream_t)stream);
}

}  // extern "C"
