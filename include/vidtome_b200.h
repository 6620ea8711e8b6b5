/*
 * vidtome_b200 — C-ABI of the B200-native VidToMe cross-frame token-merge hot path.
 *
 * The reference (lixirui142/VidToMe) is pure Python/PyTorch and has no FFI of its own; every entry
 * point below replaces a span of torch calls in the reference and cites it (file:line relative to the
 * reference tree).  The Python host side (vidtome_b200/merge.py, patch.py) binds these through
 * ctypes exactly as INTEGRATION.md shows; nothing here takes or returns a torch type.
 *
 * Conventions (all entry points):
 *   - every pointer named *_dev is a device pointer (tensor.data_ptr()); `stream` is a cudaStream_t
 *     passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - token matrices are fp16, row-major, rows of C contiguous halfs, C % 8 == 0, 16-byte aligned;
 *   - index maps are int32; "batch stride 0" means one map shared by all samples (align_batch);
 *   - the functions never allocate, never synchronise the stream, and never throw; the only process-wide
 *     state is a handful of write-once memoised lookups (driver entry point, SM count, per-kernel shared-memory opt-in
 *     and occupancy), immutable after first use, so calls are re-entrant across streams and devices;
 *   - return value: 0 = ok, <0 = bad argument (VTM_E_*), >0 = a cudaError_t / CUresult reported by
 *     the launch.  There is no CPU fallback: without a CUDA device the compute entry points return
 *     the CUDA error of the failed launch.
 */
#ifndef VIDTOME_B200_H_
#define VIDTOME_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VTM_VERSION 100 /* 0.1.0 */

#define VTM_OK 0
#define VTM_E_NULL (-1)   /* required pointer is NULL */
#define VTM_E_SHAPE (-2)  /* size / divisibility constraint violated */
#define VTM_E_SPLIT (-3)  /* inconsistent vtm_split_t */
#define VTM_E_WS (-4)     /* workspace too small */
#define VTM_E_DRIVER (-5) /* CUDA driver entry point (cuTensorMapEncodeTiled) unavailable */
#define VTM_E_UNSUPPORTED (-6) /* configuration not supported by this build of the kernel */

/*
 * How one matching level partitions its token sequence (length N) into src and dst.
 *   mode 0 (local, random target frame) — bipartite_soft_matching_randframe, vidtome/merge.py:41-74:
 *     sequence = [unm_pre carried tokens | F frames of tnum tokens]; a frame f is dst iff
 *     f % stride == randf; dst = those frames (ascending) followed by the unm_pre carried tokens;
 *     src = the remaining frames (ascending).  stride = min(target_stride, F) (merge.py:55),
 *     randf is the torch.randint value (merge.py:56-57): either drawn to the host (`randf`) or left on the
 *     device (`randf_dev` != NULL: the kernels read *randf_dev, and `randf` is ignored).  The device form lets
 *     a whole denoising step be captured in a CUDA graph whose replays draw fresh values; it requires
 *     F % stride == 0, so that the src / dst counts do not depend on the draw (else VTM_E_SPLIT).
 *   mode 1 (global, prefix/suffix) — bipartite_soft_matching_2s, vidtome/merge.py:371-379:
 *     src = positions [0, src_len), dst = positions [src_len, N).
 */
typedef struct vtm_split {
  int32_t mode;
  int32_t N;
  int32_t unm_pre;
  int32_t F;
  int32_t tnum;
  int32_t stride;
  int32_t randf;
  int32_t src_len;
  const int32_t* randf_dev; /* optional device-resident randf (mode 0), NULL = use `randf` */
} vtm_split_t;

/* Library version (VTM_VERSION).  Pure host. */
int vtm_version(void);

/* Human-readable text for a return code of this library (negative codes) or "cuda error". */
const char* vtm_error_string(int code);

/* Host helper: number of src / dst tokens of a split (merge.py:63-71 `a_idx`, `b_idx`, `num_dst`;
 * merge.py:375-379 for mode 1).  Pure host, no CUDA. */
int vtm_split_counts(const vtm_split_t* split, int32_t* num_src, int32_t* num_dst);

/* Host helper: r = min(Ns, int(Ns * ratio)) with Python-double truncation (merge.py:90, :395). */
int32_t vtm_merge_count(int32_t num_src, double ratio);

/*
 * K0 — row-normalise and split.  Replaces `metric / metric.norm(dim=-1, keepdim=True)` followed by
 * `split(metric)` (two torch.gather calls), vidtome/merge.py:76-85 and :381-390.
 *   x_dev        [B, *, C] fp16 original tokens; sample stride x_batch_stride elements
 *   rowmap_dev   [B'|1, split->N] int32: position in this level's sequence -> row of x (NULL = identity);
 *                rowmap_batch_stride = 0 shares one map across samples
 *   a_out_dev    [B, Ns, C] fp16 normalised src rows;  b_out_dev [B, Nd, C] fp16 normalised dst rows
 * Normalisation follows torch half semantics: fp32 sum of squares -> sqrt -> round to fp16 -> each
 * element fp16(float(x) / float(norm)); no epsilon (merge.py:84).
 */
int vtm_normalize_split(const void* x_dev, int64_t x_batch_stride, const int32_t* rowmap_dev,
                        int64_t rowmap_batch_stride, const vtm_split_t* split, int32_t B, int32_t C,
                        void* a_out_dev, void* b_out_dev, void* stream);

/*
 * K0 with the block's LayerNorm fused in front: rows are read from the RAW hidden states, normalised with
 * (ln_weight, ln_bias, ln_eps) exactly as `self.norm1(hidden_states)` does (vidtome/patch.py:146; torch
 * half semantics: fp32 statistics, y = gamma * (rstd * (x - mean)) + beta rounded to fp16) and then handled
 * as in vtm_normalize_split.  norm1's output is never written to memory.  ln_weight_dev == NULL disables
 * the LayerNorm (== vtm_normalize_split); ln_bias_dev may be NULL.
 */
int vtm_normalize_split_ln(const void* x_dev, int64_t x_batch_stride, const int32_t* rowmap_dev,
                           int64_t rowmap_batch_stride, const vtm_split_t* split, int32_t B, int32_t C,
                           const void* ln_weight_dev, const void* ln_bias_dev, float ln_eps, void* a_out_dev,
                           void* b_out_dev, void* stream);

/*
 * KA — fused similarity + row arg-max on tcgen05 tensor cores.  Replaces
 * `scores = a @ b.transpose(-1, -2)` and `scores.max(dim=-1)` (vidtome/merge.py:87,112 and :392,416)
 * and, with align_batch != 0, `torch.cat([*scores], dim=-1)` + max (merge.py:93-97, :398-401).
 * The Ns x Nd score matrix is never written: accumulators live in TMEM, are rounded to fp16 exactly
 * as the reference's fp16 `scores` tensor is, and reduced per row with first-index tie breaking.
 *   a_dev [B, Ns, C], b_dev [B, Nd, C] fp16 (outputs of vtm_normalize_split)
 *   keys_out_dev [B' , Ns] uint64, B' = 1 if align_batch else B.  Packed result per src row:
 *       bits 32..47  order-preserving image of the fp16 row maximum (vtm_key_score())
 *       bits  0..31  0xFFFFFFFF - arg, arg = dst index j, or b*Nd + j when align_batch
 *   The buffer is zeroed by this call (cudaMemsetAsync on `stream`) before the kernel runs.
 */
int vtm_sim_argmax(const void* a_dev, const void* b_dev, int32_t B, int32_t Ns, int32_t Nd,
                   int32_t C, int32_t align_batch, uint64_t* keys_out_dev, void* stream);

/* KA built on CTA pairs (tcgen05 cta_group::2: 256-row src blocks, each CTA loads half of every dst tile).  Same
 * results bit for bit; measured and not faster on B200, kept for A/B measurements and tests. */
int vtm_sim_argmax_pair(const void* a_dev, const void* b_dev, int32_t B, int32_t Ns, int32_t Nd,
                        int32_t C, int32_t align_batch, uint64_t* keys_out_dev, void* stream);

/* Debug/verification twin of KA on CUDA cores (one thread per src row, sequential fp32 FMA over K
 * in index order).  Same outputs; used by tests to cross-check KA at sizes the CPU oracle cannot
 * reach.  Not used by the product path. */
int vtm_sim_argmax_simt(const void* a_dev, const void* b_dev, int32_t B, int32_t Ns, int32_t Nd,
                        int32_t C, int32_t align_batch, uint64_t* keys_out_dev, void* stream);

/* Unpack helpers for KA keys (pure host; the same bit rules are used on the device). */
uint16_t vtm_key_score_half_bits(uint64_t key); /* fp16 bit pattern of the row maximum */
uint32_t vtm_key_arg(uint64_t key);             /* arg as defined above */

/*
 * KB1 — stable descending order of the row maxima.  Replaces `node_max.argsort(dim=-1,
 * descending=True)` (vidtome/merge.py:98,113,402,417) with stable semantics (ties -> lower src row
 * first), which is what torch's CUDA radix path produces and what the oracle pins.
 *   keys_dev  [Bp, Ns] uint64 from KA
 *   edge_dev  [Bp, Ns] int32: edge[k] = src row with the k-th largest maximum
 *   rank_dev  [Bp, Ns] int32: rank[edge[k]] = k
 *   ws_dev    workspace of at least vtm_topr_workspace_bytes(Bp, Ns) bytes
 */
size_t vtm_topr_workspace_bytes(int32_t Bp, int32_t Ns);
int vtm_topr_sort(const uint64_t* keys_dev, int32_t Bp, int32_t Ns, int32_t* edge_dev,
                  int32_t* rank_dev, void* ws_dev, size_t ws_bytes, void* stream);

/*
 * KB2 — apply one level's match to the composed index maps.  Replaces the index bookkeeping of the
 * merge / unmerge closures (vidtome/merge.py:115-155, :419-459) and their composition through
 * func_warper (vidtome/patch.py:84-85): in "replace" mode every level is a row gather, so the whole
 * merge is one gather `merged[i] = x[mu[i]]` and the whole unmerge is one gather `out[p] = y[pi[p]]`.
 *   split, r, Ns, Nd      this level (r = vtm_merge_count)
 *   keys_dev/edge_dev/rank_dev [Bp, Ns] from KA / KB1
 *   mu_in_dev  [Bp, split->N] int32 position -> row of the level-0 token table (NULL = identity)
 *   pi_in_dev  [Bp, N0] int32 level-0 position -> position in this level's sequence (NULL = identity,
 *              requires N0 == split->N); pi_offset is added to every pi_in entry first (global stage:
 *              the local tokens sit at [pi_offset, pi_offset + L) of the concatenated sequence)
 *   mu_out_dev [Bp, (Ns - r) + Nd] int32;  pi_out_dev [Bp, N0] int32
 * Merged sequence order is [unmerged src (edge[r:]) | dst], as merge.py:133 `cat([unm, dst])`.
 */
int vtm_compose_maps(const vtm_split_t* split, int32_t r, int32_t Ns, int32_t Nd, int32_t Bp,
                     const uint64_t* keys_dev, const int32_t* edge_dev, const int32_t* rank_dev,
                     const int32_t* mu_in_dev, const int32_t* pi_in_dev, int32_t pi_offset, int32_t N0,
                     int32_t* mu_out_dev, int32_t* pi_out_dev, void* stream);

/* Decode KA/KB1 results into the reference's own index tensors (int64, as torch returns them):
 * unm_idx = edge[r:], src_idx = edge[:r], dst_idx = node_idx[src_idx] (% Nd when align_batch)
 * (vidtome/merge.py:100-108,115-117).  node_max_dev (fp16 bits) / node_idx_dev may be NULL. */
int vtm_decode_match(const uint64_t* keys_dev, const int32_t* edge_dev, int32_t Bp, int32_t Ns,
                     int32_t Nd, int32_t r, int64_t* unm_idx_dev, int64_t* src_idx_dev,
                     int64_t* dst_idx_dev, uint16_t* node_max_dev, int64_t* node_idx_dev, void* stream);

/*
 * KC — merge gather: y[b, i, :] = x[b, map[b, i], :].  Replaces merge() of every level
 * (vidtome/merge.py:119-133, :423-437) composed by func_warper (patch.py:84), in one pass.
 */
int vtm_gather_rows(const void* x_dev, int64_t x_batch_stride, const int32_t* map_dev,
                    int64_t map_batch_stride, int32_t B, int32_t L, int32_t C, void* y_dev,
                    int64_t y_batch_stride, void* stream);

/*
 * KC as the producer of the global-token exchange (the all-gather variant of vidtome/patch.py:59-82
 * sharded one chunk per GPU): same as vtm_gather_rows_ln, and every output row is ALSO stored to n_peers (<= 8)
 * further destinations `peer_y_devs[i]` — buffers of the same layout in other GPUs' memory, peer-mapped into this
 * process (NVLink P2P; e.g. torch symmetric memory).  The merged tokens reach all ranks in the pass that creates
 * them; the caller only adds a barrier before reading the peers' contributions.  peer_y_devs is a HOST array.
 */
int vtm_gather_rows_peers(const void* x_dev, int64_t x_batch_stride, const int32_t* map_dev,
                          int64_t map_batch_stride, int32_t B, int32_t L, int32_t C, const void* ln_weight_dev,
                          const void* ln_bias_dev, float ln_eps, void* y_dev, int64_t y_batch_stride,
                          void* const* peer_y_devs, int32_t n_peers, void* stream);

/* KC with norm1 fused: y[b, i, :] = LayerNorm(x[b, map[b, i], :]) — the merged tokens that feed attn1,
 * computed from the raw hidden states (only the L kept rows are ever normalised). */
int vtm_gather_rows_ln(const void* x_dev, int64_t x_batch_stride, const int32_t* map_dev,
                       int64_t map_batch_stride, int32_t B, int32_t L, int32_t C, const void* ln_weight_dev,
                       const void* ln_bias_dev, float ln_eps, void* y_dev, int64_t y_batch_stride, void* stream);

/*
 * Merge with a reduction — the non-"replace" modes of the merge closure, vidtome/merge.py:126-131:
 * `dst.scatter_reduce(-2, dst_idx, src[src_idx], reduce=mode, include_self=True)` followed by `cat([unm, dst])`.
 *   mode: 1 = "mean", 2 = "sum", 3 = "amax", 4 = "amin" ("prod" is not implemented: VTM_E_UNSUPPORTED)
 *   x_dev [B, split->N, C] fp16 (the tensor being merged; finite values), keys/edge [Bp, Ns] from KA / KB1, r as in KB2
 *   y_dev [B, (Ns - r) + Nd, C] fp16;  ws_dev: vtm_merge_reduce_workspace_bytes(B, Nd, C) bytes
 * Sums are exact and order independent (int64 fixed-point accumulators, value * 2^24) and rounded once; the mean is
 * fp16(float(sum) / float(count)) as torch computes it.  The reference accumulates in fp16 (CPU: index order; CUDA:
 * atomics in arbitrary order) and therefore agrees bit for bit only where its partial sums are exact.
 * No caller in the reference passes a mode; a reduction is not a row gather, so this is an operator-API entry and
 * not part of the composed per-block plan.
 */
size_t vtm_merge_reduce_workspace_bytes(int32_t B, int32_t Nd, int32_t C);
int vtm_merge_reduce(const void* x_dev, int64_t x_batch_stride, const vtm_split_t* split, int32_t r, int32_t Bp,
                     const uint64_t* keys_dev, const int32_t* edge_dev, int32_t B, int32_t C, int32_t mode,
                     void* y_dev, void* ws_dev, size_t ws_bytes, void* stream);

/*
 * KE — unmerge gather fused with the residual add: out[b, p, :] = y[b, map[b, p], :] + resid[b, p, :].
 * Replaces unmerge() of every level (zeros + 3 scatter_, vidtome/merge.py:135-155, :439-460),
 * split_frame (vidtome/utils.py:37-40) and `attn_output + hidden_states` (vidtome/patch.py:168-169).
 * resid_dev may be NULL (plain unmerge, used by the merge.py-level API).
 */
int vtm_unmerge_add(const void* y_dev, int64_t y_batch_stride, const int32_t* map_dev,
                    int64_t map_batch_stride, const void* resid_dev, int32_t B, int32_t N, int32_t C,
                    void* out_dev, void* stream);

/*
 * KD — merged-token self-attention.  Replaces `self.attn1(merged_tokens)` (vidtome/patch.py:157-162;
 * math restated in the reference at utils/pnp_utils.py:47-95): q,k,v = x Wq^T, x Wk^T, x Wv^T (no bias),
 * per-head softmax(q k^T * scale) v, heads re-joined, y = o Wo^T + bo.
 *   x_dev [B, L, C] fp16; w_qkv_dev [3C, C] fp16 (rows: Wq | Wk | Wv, torch Linear layout);
 *   w_o_dev [C, C] fp16; b_o_dev [C] fp16 (may be NULL); heads * head_dim == C; head_dim % 8 == 0 and head_dim <= 128 (else VTM_E_SHAPE / VTM_E_UNSUPPORTED)
 *   y_dev [B, L, C] fp16; ws_dev workspace of vtm_attention_workspace_bytes(B, L, C, heads) bytes (q/k/v head-major
 *   with rows padded to 64 or 128 halfs, plus the attention output before the out projection).
 */
size_t vtm_attention_workspace_bytes(int32_t B, int32_t L, int32_t C, int32_t heads);
int vtm_attention(const void* x_dev, const void* w_qkv_dev, const void* w_o_dev, const void* b_o_dev,
                  int32_t B, int32_t L, int32_t C, int32_t heads, float scale, void* y_dev,
                  void* ws_dev, size_t ws_bytes, void* stream);

/* KD with options.  flags bit 0 (VTM_ATTN_SHARED_QK): PnP's source-sample injection (utils/pnp_utils.py:57-68,87-91: the
 * attention map softmax(q k^T) is computed from sample 0 only and `repeat`ed over the batch): every sample b uses the
 * queries and keys of sample 0 and its own values.  Other bits must be zero (VTM_E_UNSUPPORTED). */
#define VTM_ATTN_SHARED_QK 1
int vtm_attention_ex(const void* x_dev, const void* w_qkv_dev, const void* w_o_dev, const void* b_o_dev,
                     int32_t B, int32_t L, int32_t C, int32_t heads, float scale, int32_t flags, void* y_dev,
                     void* ws_dev, size_t ws_bytes, void* stream);

/* Plain tcgen05 GEMM used by KD's projections, exported for tests and the microbench:
 * D[M, N] = A[M, K] * W[N, K]^T (+ bias[N]), fp16 in, fp32 accumulate, fp16 out.  K % 8 == 0,
 * N % 8 == 0; ldd = row stride of D in elements. */
int vtm_linear_f16(const void* a_dev, const void* w_dev, const void* bias_dev, int32_t M, int32_t N,
                   int32_t K, void* d_dev, int64_t ldd, void* stream);

/*
 * Cross-attention of the patched block.  Replaces `self.attn2(norm2(h), encoder_hidden_states=ctx) + h`
 * (vidtome/patch.py:171-185; diffusers Attention with K/V taken from the text context):
 *   q = x Wq^T, [k | v] = ctx [Wk; Wv]^T (no bias), per-head softmax(q k^T * scale) v, y = o Wo^T + bo (+ resid).
 *   x_dev [B, Lq, C] fp16 (norm2 output); ctx_dev [B, Lk, Cctx] fp16; w_q_dev [C, C]; w_kv_dev [2C, Cctx] (rows Wk | Wv);
 *   w_o_dev [C, C]; b_o_dev [C] or NULL; resid_dev [B, Lq, C] or NULL; y_dev [B, Lq, C].
 *   B counts (sample x frame) items: every item attends to its own Lk context rows.  head_dim % 8 == 0, <= 128.
 */
size_t vtm_cross_attention_workspace_bytes(int32_t B, int32_t Lq, int32_t Lk, int32_t C, int32_t heads);
int vtm_cross_attention(const void* x_dev, const void* ctx_dev, const void* w_q_dev, const void* w_kv_dev,
                        const void* w_o_dev, const void* b_o_dev, const void* resid_dev, int32_t B, int32_t Lq,
                        int32_t Lk, int32_t C, int32_t Cctx, int32_t heads, float scale, void* y_dev, void* ws_dev,
                        size_t ws_bytes, void* stream);

/*
 * Feed-forward of the patched block (vidtome/patch.py:187-199 calls `self.ff(norm3(h))` and adds the residual; the
 * module is diffusers' FeedForward = GEGLU(dim -> 4 dim) -> Dropout -> Linear(4 dim -> dim)) on tcgen05:
 *
 * vtm_linear_geglu_f16:  D[M, N/2] = h * gelu(gate) with [h | gate] = A[M, K] * W^T + bias (erf form of gelu, torch's
 *   F.gelu default).  `w_il_dev` [N, K] and `bias_il_dev` [N] hold the projection's rows INTERLEAVED in groups of 32:
 *   rows [64 q, 64 q + 32) = rows [32 q, 32 q + 32) of the value half, rows [64 q + 32, 64 q + 64) = the same rows of the
 *   gate half (so that a value and its gate land in one 64-column TMEM load).  N % 64 == 0.  Roundings follow torch's
 *   fp16 pipeline: projection -> fp16, gelu(gate) -> fp16, product -> fp16.
 * vtm_linear_residual_f16:  D[M, N] = fp16(fp16(A W^T + bias) + resid[M, N]) — `lin(x) + hidden_states`.
 */
int vtm_linear_geglu_f16(const void* a_dev, const void* w_il_dev, const void* bias_il_dev, int32_t M, int32_t N,
                         int32_t K, void* d_dev, int64_t ldd, void* stream);
int vtm_linear_residual_f16(const void* a_dev, const void* w_dev, const void* bias_dev, const void* resid_dev,
                            int64_t ldr, int32_t M, int32_t N, int32_t K, void* d_dev, int64_t ldd, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIDTOME_B200_H_ */
