// Reranker hand-off on the device: the coarse ranker's fused top-k ids -> the token sequences
// the LLM reranker scores, in the reference's layout and slice order, without a round trip through Python lists.
//
// Reference: LLMRerank._postprocess_nodes walks the coarse list in slices of embed_bs = 32 (rerankers.py:309-322)
// and get_inputs / get_inputs_v2_5 (rerankers.py:196-293) builds, per (query, passage) pair,
//     [bos] + tok("A: " + query)[: 3/4 max_length]  |  sep + tok("B: " + passage)   (the pair truncated to max_length,
//     'only_second': the passage side gives way)     |  sep + prompt
// Here the passages ("B: ..." already tokenised once at index time, CSR on the device), the queries ("A: ...",
// CSR per batch) and the sep / prompt ids are device arrays; pair p = q * k + r is candidate r of query q.
// Output is PACKED (no padding): ids[T] + cu_seqlens[P + 1] (the layout this library's encoder kernels consume;
// a padded [32, L] view is a copy with cu as the index), plus the per-pair query_lengths of get_inputs_v2_5.
#include "ezr_common.cuh"
#include "../../include/easyrag_b200.h"

namespace ezr {

struct RerankParams {
    const int32_t* cand_ids;    // [Q, k_stride] document ids in rank order (-1 padded)
    const int32_t* cand_cnt;    // [Q]
    int n_queries, k, k_stride, id_base;
    const int32_t* q_ptr;       // [Q + 1] into q_tok
    const int32_t* q_tok;       // query tokens WITHOUT bos
    const int64_t* p_ptr;       // [N + 1] into p_tok
    const int32_t* p_tok;
    const int32_t* sep;         // [n_sep]
    const int32_t* prompt;      // [n_prompt]
    int n_sep, n_prompt, bos, max_length;
};

// tokens of the three parts of pair p: head = bos + query (<= 3/4 max_length query tokens), body = sep + passage
// truncated so that head + body <= max_length, tail = sep + prompt
__device__ __forceinline__ bool pair_parts(const RerankParams& a, int p, int& doc, int& nq, int& npass) {
    const int q = p / a.k, r = p % a.k;
    doc = -1; nq = 0; npass = 0;
    if (r >= a.cand_cnt[q]) return false;
    doc = a.cand_ids[(int64_t)q * a.k_stride + r] - a.id_base;
    if (doc < 0) return false;
    nq = min(a.q_ptr[q + 1] - a.q_ptr[q], a.max_length * 3 / 4);
    const int64_t plen = a.p_ptr[doc + 1] - a.p_ptr[doc];
    const int head = 1 + nq;
    const int room = a.max_length - head - a.n_sep;              // 'only_second': the passage gives way
    npass = (int)min((int64_t)max(room, 0), min(plen, (int64_t)a.max_length));
    return true;
}

__global__ void rerank_len_kernel(const RerankParams a, int64_t* __restrict__ len, int32_t* __restrict__ query_len) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.n_queries * a.k) return;
    int doc, nq, npass;
    if (!pair_parts(a, p, doc, nq, npass)) { len[p] = 0; query_len[p] = 0; return; }
    len[p] = 1 + nq + a.n_sep + npass + a.n_sep + a.n_prompt;
    query_len[p] = 1 + nq + a.n_sep;                             // get_inputs_v2_5: len([bos] + query + sep)
}

// one warp per pair
__global__ void rerank_fill_kernel(const RerankParams a, const int64_t* __restrict__ cu, int32_t* __restrict__ ids,
                                   int32_t* __restrict__ cu32) {
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int n_pairs = a.n_queries * a.k;
    if (p > n_pairs) return;
    if (lane == 0) cu32[p] = (int32_t)cu[p];
    if (p == n_pairs) return;
    int doc, nq, npass;
    if (!pair_parts(a, p, doc, nq, npass)) return;
    const int q = p / a.k;
    int32_t* o = ids + cu[p];
    if (lane == 0) o[0] = a.bos;
    const int32_t* qs = a.q_tok + a.q_ptr[q];
    for (int i = lane; i < nq; i += 32) o[1 + i] = qs[i];
    o += 1 + nq;
    for (int i = lane; i < a.n_sep; i += 32) o[i] = a.sep[i];
    o += a.n_sep;
    const int32_t* ps = a.p_tok + a.p_ptr[doc];
    for (int i = lane; i < npass; i += 32) o[i] = ps[i];
    o += npass;
    for (int i = lane; i < a.n_sep; i += 32) o[i] = a.sep[i];
    o += a.n_sep;
    for (int i = lane; i < a.n_prompt; i += 32) o[i] = a.prompt[i];
}

// exclusive scan of int64 lengths (one CTA; P is at most a few hundred thousand pairs)
__global__ void __launch_bounds__(1024)
rerank_scan_kernel(const int64_t* __restrict__ len, int n, int64_t* __restrict__ cu) {
    __shared__ long long s_warp[32];
    __shared__ long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_carry = 0; cu[0] = 0; }
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        long long v = i < n ? (long long)len[i] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        if (lane == 31) s_warp[warp] = v;
        __syncthreads();
        if (warp == 0) {
            long long w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += u;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const long long incl = v + (warp > 0 ? s_warp[warp - 1] : 0) + s_carry;
        if (i < n) cu[i + 1] = incl;
        __syncthreads();
        if (tid == 1023) s_carry = incl;
        __syncthreads();
    }
}

static int fill_params(RerankParams& a, const int32_t* cand_ids, const int32_t* cand_cnt, int n_queries, int k, int k_stride,
                       int id_base, const int32_t* q_ptr, const int32_t* q_tok, const int64_t* p_ptr, const int32_t* p_tok,
                       const int32_t* sep, int n_sep, const int32_t* prompt, int n_prompt, int bos, int max_length) {
    EZR_CHECK_ARG(cand_ids && cand_cnt && q_ptr && p_ptr && (q_tok || true) && p_tok, "rerank_pack: NULL argument");
    EZR_CHECK_ARG(n_queries >= 0 && k >= 1 && k_stride >= k, "rerank_pack: bad n_queries / k / stride");
    EZR_CHECK_ARG(n_sep >= 0 && n_prompt >= 0 && (n_sep == 0 || sep) && (n_prompt == 0 || prompt) && max_length >= 8,
                  "rerank_pack: bad sep / prompt / max_length");
    EZR_CHECK_ARG((int64_t)n_queries * k < ((int64_t)1 << 30), "rerank_pack: too many pairs");
    EZR_CHECK_ARG(1 + max_length * 3 / 4 + n_sep <= max_length,
                  "rerank_pack: max_length=%d leaves no room for bos + query + sep (%d sep ids)", max_length, n_sep);
    a.cand_ids = cand_ids; a.cand_cnt = cand_cnt; a.n_queries = n_queries; a.k = k; a.k_stride = k_stride;
    a.id_base = id_base; a.q_ptr = q_ptr; a.q_tok = q_tok; a.p_ptr = p_ptr; a.p_tok = p_tok; a.sep = sep; a.prompt = prompt;
    a.n_sep = n_sep; a.n_prompt = n_prompt; a.bos = bos; a.max_length = max_length;
    return EZR_OK;
}

}  // namespace ezr

using namespace ezr;

extern "C" {

int ezr_rerank_pack_plan(const int32_t* cand_ids, const int32_t* cand_cnt, int32_t n_queries, int32_t k, int32_t k_stride,
                         int32_t id_base, const int32_t* q_ptr, const int64_t* p_ptr, int32_t n_sep, int32_t n_prompt,
                         int32_t max_length, int64_t* out_len, int64_t* out_cu, int32_t* out_query_len,
                         int64_t* total_host, void* stream) {
    RerankParams a;
    int rc = fill_params(a, cand_ids, cand_cnt, n_queries, k, k_stride, id_base, q_ptr, nullptr, p_ptr,
                         reinterpret_cast<const int32_t*>(1), n_sep ? reinterpret_cast<const int32_t*>(1) : nullptr, n_sep,
                         n_prompt ? reinterpret_cast<const int32_t*>(1) : nullptr, n_prompt, 0, max_length);
    if (rc) return rc;
    EZR_CHECK_ARG(out_len && out_cu && out_query_len && total_host, "rerank_pack_plan: NULL output");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_pairs = n_queries * k;
    if (n_pairs == 0) { *total_host = 0; return EZR_OK; }
    rerank_len_kernel<<<ceil_div(n_pairs, 256), 256, 0, st>>>(a, out_len, out_query_len);
    EZR_LAUNCH_CHECK();
    rerank_scan_kernel<<<1, 1024, 0, st>>>(out_len, n_pairs, out_cu);
    EZR_LAUNCH_CHECK();
    EZR_CUDA(cudaMemcpyAsync(total_host, out_cu + n_pairs, 8, cudaMemcpyDeviceToHost, st));
    EZR_CUDA(cudaStreamSynchronize(st));
    return EZR_OK;
}

int ezr_rerank_pack_fill(const int32_t* cand_ids, const int32_t* cand_cnt, int32_t n_queries, int32_t k, int32_t k_stride,
                         int32_t id_base, const int32_t* q_ptr, const int32_t* q_tok, const int64_t* p_ptr,
                         const int32_t* p_tok, const int32_t* sep, int32_t n_sep, const int32_t* prompt, int32_t n_prompt,
                         int32_t bos, int32_t max_length, const int64_t* cu, int32_t* out_ids, int32_t* out_cu32,
                         void* stream) {
    RerankParams a;
    int rc = fill_params(a, cand_ids, cand_cnt, n_queries, k, k_stride, id_base, q_ptr, q_tok, p_ptr, p_tok, sep, n_sep,
                         prompt, n_prompt, bos, max_length);
    if (rc) return rc;
    EZR_CHECK_ARG(cu && out_ids && out_cu32 && q_tok, "rerank_pack_fill: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_pairs = n_queries * k;
    rerank_fill_kernel<<<ceil_div(n_pairs + 1, 8), 256, 0, st>>>(a, cu, out_ids, out_cu32);
    EZR_LAUNCH_CHECK();
    return EZR_OK;
}

}  // extern "C"
