/*
 * rf_b200.h -- C-ABI of the B200-native feature-to-embedding hot path of RecommendFlow.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / TF types.  Every entry
 * point is stream-ordered (pass a cudaStream_t as `void*`, NULL = legacy default stream),
 * returns RF_OK (0) or a negative rf_status, and leaves a message for rf_last_error()
 * (thread-local) on failure.  There is NO CPU fallback: without a CUDA device every compute
 * entry point fails with RF_ERR_CUDA.
 *
 * The reference (/root/reference, pure Python on TensorFlow/Keras ops) has no FFI of its own;
 * each entry point below names the reference interface it replaces.  INTEGRATION.md shows
 * the ctypes / TF-custom-op stubs a maintainer of the reference would add.
 */
#ifndef RF_B200_H_
#define RF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RF_B200_ABI_VERSION 3

typedef enum rf_status {
    RF_OK = 0,
    RF_ERR_INVALID = -1,     /* bad argument (the Python layer raises ValueError)            */
    RF_ERR_CUDA = -2,        /* CUDA runtime error or no device                               */
    RF_ERR_UNSUPPORTED = -3  /* valid in the reference but outside this library's envelope    */
} rf_status;

/* Combiners of EmbeddingBag.get_combiner (backend/layers/preprocess_layers.py:43-64).
 * "null"/"first"/"last" are resolved by the host layer as bags of one item (a plain gather). */
typedef enum rf_combiner {
    RF_COMBINER_SUM = 0,
    RF_COMBINER_AVG = 1,
    RF_COMBINER_MIN = 2,
    RF_COMBINER_MAX = 3
} rf_combiner;

/* Keras `Hashing(mask_value=...)`: which input value is sent to bucket 0. */
typedef enum rf_mask_mode {
    RF_MASK_NONE = 0,        /* mask_value=None: ids in [0, num_bins)                          */
    RF_MASK_EMPTY_STRING = 1,/* mask_value="" (what get_preprocess_layers passes,             */
                             /* backend/utils/preprocess_utils.py:15): "" -> 0, else 1+h%(N-1) */
    RF_MASK_INT_VALUE = 2,   /* integer inputs: value == int_mask_value -> 0                  */
    RF_MASK_STRING_VALUE = 3 /* mask_value="<any string>": keys whose bytes equal                  */
                             /* rf_field_desc.mask_bytes[0 .. mask_len) -> 0 (Keras compares the   */
                             /* raw strings before hashing); at most RF_MAX_MASK_BYTES bytes        */
} rf_mask_mode;
#define RF_MAX_MASK_BYTES 32

/* One embedding table + the hash that feeds it: replaces one Keras `Hashing` + `Embedding`
 * pair (preprocess_layers.py:89-92, :31-39). */
typedef struct rf_table_desc {
    const float *weights;    /* device, [num_bins, dim] fp32 row-major (Embedding.embeddings)  */
    int64_t num_bins;        /* Hashing num_bins == Embedding input_dim                        */
    int32_t use_strong;      /* 1: salt given -> SipHash-2-4 (to_hash_bucket_strong)           */
                             /* 0: salt None  -> FarmHash Fingerprint64 (to_hash_bucket_fast)  */
    int32_t reserved;
    uint64_t key0, key1;     /* SipHash key: salt=[key0,key1]; int salt s -> (s, s)            */
} rf_table_desc;

#define RF_MAX_TABLES_PER_FIELD 2
/* rf_field_desc.flags: this launch produces PARTIAL pools (row-sharded tables): an empty bag    */
/* yields the combiner's identity (+/-inf for min/max) instead of 0.                             */
#define RF_FIELD_PARTIAL 1
/* The pooled vector is ADDED into `out` (red.global.add, also over NVLink peer pointers) instead of  */
/* stored; empty bags add nothing.  sum / avg only; must be set on every field of a launch or none.  */
/* `out` must be zeroed (or hold the running value) before the launch.  Summation order across       */
/* concurrent launches / owners is free, i.e. results are not bit-reproducible.                       */
#define RF_FIELD_ACCUMULATE 2

/* One feature field of one batch: replaces `DoubleHashingEmbedding.call`
 * (preprocess_layers.py:94-97; n_tables == 2) or `EmbeddingBag.call` (:66-68; pre-hashed
 * ids, n_tables == 1).  Exactly one of {bytes+str_offsets, int_values, ids} is non-NULL. */
typedef struct rf_field_desc {
    /* --- input keys, flat over the field's items (dense: item = b * bag_len + l) ----------- */
    const uint8_t *bytes;        /* device string arena; >= 16 readable bytes past the end     */
    const int32_t *str_offsets;  /* device [n_items + 1], byte offsets into `bytes`            */
    const int64_t *int_values;   /* device [n_items]; hashed as tf.as_string(value)            */
    const int64_t *ids;          /* device [n_tables][n_items] pre-hashed row ids (no hashing) */
    /* --- bags -------------------------------------------------------------------------------- */
    const int32_t *bag_offsets;  /* device [batch + 1] CSR over items (jagged mode), or NULL   */
    const int32_t *bag_ends;     /* optional device [batch]: bag b = [bag_offsets[b],          */
                                 /* bag_ends[b]) -- bags in order but with gaps between them   */
                                 /* (then bag_offsets needs only [batch] entries)              */
    int64_t n_items;             /* jagged mode: total items (= bag_offsets[batch]); dense     */
                                 /* mode derives batch * bag_len and ignores this              */
    int32_t bag_len;             /* dense mode (bag_offsets == NULL): items per bag, pads      */
                                 /* included -- the reference pools row 0 in for every pad     */
    int32_t n_tables;            /* 1 .. RF_MAX_TABLES_PER_FIELD                                */
    rf_table_desc tables[RF_MAX_TABLES_PER_FIELD];
    int32_t dim;                 /* embedding dim D; 0 = hash only (needs ids_out)             */
    int32_t combiner;            /* rf_combiner                                                 */
    int32_t mask_mode;           /* rf_mask_mode                                                */
    int32_t flags;               /* RF_FIELD_* bits                                             */
    int64_t int_mask_value;      /* RF_MASK_INT_VALUE only                                      */
    /* --- outputs ----------------------------------------------------------------------------- */
    float *out;                  /* device; bag b, table t -> out[b*out_stride + t*dim .. +dim] */
    int64_t out_stride;          /* floats between consecutive bags' rows                       */
    int64_t *ids_out;            /* optional device [n_tables][n_items]: the bucket ids         */
    /* --- RF_MASK_STRING_VALUE only (host bytes, copied into the launch descriptor) ------------ */
    uint8_t mask_bytes[RF_MAX_MASK_BYTES];
    int32_t mask_len;
    int32_t reserved;
} rf_field_desc;

/* ---- library ------------------------------------------------------------------------------ */
int rf_abi_version(void);
const char *rf_last_error(void);

/* ---- hashing only: Keras `Hashing.call` (preprocess_layers.py:95) ------------------------- */
int rf_hash_strings(const uint8_t *d_bytes, const int32_t *d_str_offsets, int64_t n_items,
                    int64_t num_bins, int mask_mode, int use_strong, uint64_t key0, uint64_t key1,
                    int64_t *d_ids_out, void *stream);
int rf_hash_int64(const int64_t *d_values, int64_t n_items, int64_t num_bins, int mask_mode,
                  int64_t int_mask_value, int use_strong, uint64_t key0, uint64_t key1,
                  int64_t *d_ids_out, void *stream);
/* Keras `Hashing(num_bins, mask_value="<string>")`: keys equal to the mask string go to bucket 0. */
/* h_mask_value: HOST bytes, mask_len <= RF_MAX_MASK_BYTES (mask_len == 0 is mask_value="").       */
int rf_hash_strings_masked(const uint8_t *d_bytes, const int32_t *d_str_offsets, int64_t n_items,
                           int64_t num_bins, const uint8_t *h_mask_value, int32_t mask_len, int use_strong,
                           uint64_t key0, uint64_t key1, int64_t *d_ids_out, void *stream);

/* ---- fused hash + gather + pool over n_fields fields of one batch, ONE kernel launch ------- */
/* Replaces the per-feature loop `self.preprocessor[name](batch[name])`                        */
/* (models/matching/que2search.py:68,76-79) over the layers built by get_preprocess_layers     */
/* (backend/utils/preprocess_utils.py:7-20).                                                   */
int rf_bag_forward(const rf_field_desc *fields, int n_fields, int64_t batch, void *stream);
/* Same, with the grid capped at max_ctas_per_sm x #SMs for THIS launch (0 = no cap): the kernel then walks */
/* its tiles grid-stride and leaves room on every SM for kernels of other streams (sharded pipeline).      */
int rf_bag_forward_ex(const rf_field_desc *fields, int n_fields, int64_t batch, int max_ctas_per_sm, void *stream);
/* Launches recorded into CUDA graphs keep a descriptor slot each (256 per device).  Call this after the   */
/* graphs that hold them were destroyed to hand all slots back.                                             */
int rf_release_captured_launches(void);

/* ---- scaled_dot_product_attention (backend/layers/layer_utils.py:4-24), exact fp32 ------------ */
/* q, k, v, out: device [n_batch_heads, seq_len, head_dim] fp32; mask: device [n_batch_heads,    */
/* seq_len] fp32 or NULL.  mask[i] == 0 fills QUERY row i of the logits with -4294967295 (the    */
/* reference's [..., S, 1] mask broadcasts over keys); scale 1/sqrt(head_dim); softmax over keys. */
int rf_sdpa_forward(const float *d_q, const float *d_k, const float *d_v, const float *d_mask,
                    int64_t n_batch_heads, int32_t seq_len, int32_t head_dim, float *d_out, void *stream);

/* Same contract on the tensor cores (tcgen05.mma kind::tf32, TMEM accumulators, TMA): pairs of   */
/* sequences share one 128-row tile; seq_len <= 64, head_dim in {32, 64, 96}; otherwise           */
/* RF_ERR_UNSUPPORTED (use rf_sdpa_forward).  fp32 operands are read as TF32.                     */
int rf_sdpa_forward_tc(const float *d_q, const float *d_k, const float *d_v, const float *d_mask,
                       int64_t n_batch_heads, int32_t seq_len, int32_t head_dim, float *d_out, void *stream);

/* Same, with q / k / v rows `row_pitch` floats apart (out stays contiguous): q, k, v may be column windows of ONE    */
/* fused projection output [rows, 3 * head_dim] (MultiHeadAttention with a single Dense for the three projections). */
int rf_sdpa_forward_tc_strided(const float *d_q, const float *d_k, const float *d_v, int64_t row_pitch, const float *d_mask,
                               int64_t n_batch_heads, int32_t seq_len, int32_t head_dim, float *d_out, void *stream);

/* ---- in-batch two-tower logits S = query . doc^T, reduced per row without ever storing S ------ */
/* (backend/lossess/match_losses.py:119-226 all start from tf.matmul(query, tf.transpose(doc))).   */
/* Per row i (any output pointer may be NULL):                                                     */
/*   lse[i]    = log sum_j exp(scale * S_ij)       diag[i] = S_ii                                 */
/*   hinge[i]  = sum_j clip(S_ij - S_ii + margin, 0, 1e14) * (col_weight ? col_weight[j] : 1)      */
/*   maxoff[i] = max_j (j == i ? 0 : S_ij)                                                         */
/*   *loss     = mean_i( -(scale * S_ii - lse[i]) * y[i] )   = batch_neg_sample_scaled_multi_class */
/*               _ce_loss (:150-165), evaluated with max-subtraction (same value, no overflow).    */
/* d_workspace: rf_inbatch_workspace_bytes(batch) bytes of device scratch.                        */
int64_t rf_inbatch_workspace_bytes(int64_t batch);
int rf_inbatch_rowstats(const float *d_query, const float *d_doc, const float *d_y, const float *d_col_weight,
                        int64_t batch, int32_t dim, float scale, float margin, void *d_workspace,
                        float *d_lse, float *d_diag, float *d_hinge, float *d_maxoff, float *d_loss,
                        void *stream);

/* Same contract on the tensor cores: tcgen05.mma kind::tf32 (fp32 operands read as TF32, fp32   */
/* accumulators in TMEM, TMA-fed 128x256x32 tiles, reductions fused into the TMEM epilogue).     */
/* Operands are first rounded to nearest TF32 (unbiased) into the workspace; the diagonal S_ii   */
/* stays an exact fp32 dot product.  Needs dim % 4 == 0 and                                      */
/* rf_inbatch_workspace_bytes_tc(batch, dim) bytes of workspace.                                 */
int64_t rf_inbatch_workspace_bytes_tc(int64_t batch, int32_t dim);
int rf_inbatch_rowstats_tc(const float *d_query, const float *d_doc, const float *d_y, const float *d_col_weight,
                           int64_t batch, int32_t dim, float scale, float margin, void *d_workspace,
                           float *d_lse, float *d_diag, float *d_hinge, float *d_maxoff, float *d_loss,
                           void *stream);

/* ---- Keras Dense on the tensor cores: out = activation(x . W + b) (+ optional row l2-normalisation) ---------- */
/* The tower MLP of backend/blocks/mlp.py:4-15 ([norm, Dense(units, activation), Dropout] * n; models/matching/     */
/* dssm.py:25-26: [1024, 512, 256], selu, BatchNormalization(1e-6)) and the Dense q/k/v projections of              */
/* backend/layers/attention_layers.py:141-155.  x: device [rows, in_dim] fp32 with row pitch ldx (floats);           */
/* weight_t: device [units, in_dim] fp32 = the TRANSPOSE of the Keras kernel [in_dim, units] (K-major for the       */
/* tensor core); bias: device [units] or NULL; out: device [rows, units] with row pitch ldo.  Operands are read as  */
/* TF32 (fp32 accumulate).  An inference-mode BatchNormalization in front of the Dense is folded into weight_t /    */
/* bias by the caller.  l2_normalize != 0 divides every output row by max(||row||_2, 1e-12) in the same epilogue    */
/* (units <= 256).  in_dim, units, ldx, ldo multiples of 4; 16-byte aligned buffers; else RF_ERR_UNSUPPORTED.        */
typedef enum rf_activation {
    RF_ACT_NONE = 0, RF_ACT_RELU = 1, RF_ACT_SELU = 2, RF_ACT_TANH = 3, RF_ACT_SIGMOID = 4, RF_ACT_GELU = 5
} rf_activation;
int rf_dense_forward_tc(const float *d_x, int64_t rows, int32_t in_dim, int64_t ldx, const float *d_weight_t,
                        const float *d_bias, int32_t units, int activation, int l2_normalize, float *d_out,
                        int64_t ldo, void *stream);
/* The same product with a workspace: when the output has too few 128 x 64 tiles to occupy the GPU and the contraction is  */
/* long (dW = X^T dZ of a tower stage: 4 x 4 tiles, K = batch), K is split over CTAs, the partial products go to the       */
/* workspace and are summed in a fixed order (deterministic).  Only for bias == NULL, activation none, no normalisation;  */
/* otherwise, or with d_workspace == NULL / too small, identical to rf_dense_forward_tc.                                   */
int64_t rf_dense_tc_workspace_bytes(int64_t rows, int32_t in_dim, int32_t units);
int rf_dense_forward_tc_ex(const float *d_x, int64_t rows, int32_t in_dim, int64_t ldx, const float *d_weight_t,
                           const float *d_bias, int32_t units, int activation, int l2_normalize, float *d_out,
                           int64_t ldo, void *d_workspace, int64_t workspace_bytes, void *stream);

/* Same contract with bf16 operands (kind::f16, fp32 accumulate): the TF32 kernel is bound by L2 -> SM operand   */
/* traffic; here a CTA keeps 256 query rows resident and streams 2-byte doc tiles (4x less traffic per flop).     */
/* Operands are rounded to nearest bf16 into the workspace (rf_inbatch_workspace_bytes_tc is enough); the         */
/* diagonal stays exact fp32.  dim % 8 == 0 and dim <= 256, else RF_ERR_UNSUPPORTED.                               */
int rf_inbatch_rowstats_bf16(const float *d_query, const float *d_doc, const float *d_y, const float *d_col_weight,
                             int64_t batch, int32_t dim, float scale, float margin, void *d_workspace,
                             float *d_lse, float *d_diag, float *d_hinge, float *d_maxoff, float *d_loss,
                             void *stream);

/* ---- row-sharded tables (new design; the reference only replicates tables,   */
/* backend/utils/gpu_utils.py:13-14).  Row id lives on rank id % world as local row id / world. */
/* rf_shard_route partitions the hashed ids of one field by owner, keeping bag order: for every */
/* owner g it writes a CSR offsets[batch+1] through h_offsets_dst[g] (may be NULL) and the      */
/* owner-local rows through h_rows_dst[g] (int64, capacity >= this rank's key count).  The      */
/* h_* arrays are HOST arrays of `world` DEVICE pointers -- peer-mapped NVLink pointers in the  */
/* fused path, local send buffers in the NCCL path.  Scratch: d_counts_ws int32[world*batch],   */
/* d_offsets_local int32[world*(batch + 1 + ceil(batch/1024))] (in-chunk scans + chunk totals). */
int rf_shard_route(const int64_t *d_ids, const int32_t *d_bag_offsets, int32_t bag_len, int64_t batch,
                   int world, int32_t *d_counts_ws, int32_t *d_offsets_local,
                   int32_t *const *h_offsets_dst, int64_t *const *h_rows_dst, void *stream);
/* Same, fused with the hashing of string keys (no separate pass over the batch): d_ids_ws       */
/* (int64[n_keys]) receives the bucket ids.                                                      */
int rf_shard_route_keys(const uint8_t *d_bytes, const int32_t *d_str_offsets, int64_t num_bins, int mask_mode,
                        int use_strong, uint64_t key0, uint64_t key1, int64_t *d_ids_ws,
                        const int32_t *d_bag_offsets, int32_t bag_len, int64_t batch, int world,
                        int32_t *d_counts_ws, int32_t *d_offsets_local, int32_t *const *h_offsets_dst,
                        int64_t *const *h_rows_dst, void *stream);
/* Single-pass routing into the "tile" layout (no global scan, one kernel): for every owner g the */
/* owner-local rows of this source's keys k0..k1 land in [k0, k0 + n_g) of h_rows_dst[g] (int64,  */
/* capacity = this rank's key count; the rest of each range stays unused) and bag b's run is      */
/* [h_begin_dst[g][b], h_end_dst[g][b]) -- feed them to rf_bag_forward as ids / bag_offsets /     */
/* bag_ends.  Keys are strings (d_bytes + d_str_offsets, hashed on the fly; d_ids_ws receives the */
/* ids; may be NULL when the caller does not need them) or pre-hashed d_ids.                       */
int rf_shard_route_tiles(const uint8_t *d_bytes, const int32_t *d_str_offsets, const int64_t *d_ids,
                         int64_t num_bins, int mask_mode, int use_strong, uint64_t key0, uint64_t key1,
                         int64_t *d_ids_ws, const int32_t *d_bag_offsets, int32_t bag_len, int64_t batch,
                         int world, int64_t *const *h_rows_dst, int32_t *const *h_begin_dst,
                         int32_t *const *h_end_dst, void *stream);
/* Same, with the routing kernel's grid capped at max_ctas_per_sm x #SMs (0 = no cap).  In the pipelined step the  */
/* routing of batch i+1 runs under the HBM-bound pooling of batch i: 2 CTAs per SM measured best (more starve the  */
/* pooling kernel of residency, fewer leave the routing latency-bound).                                            */
int rf_shard_route_tiles_ex(const uint8_t *d_bytes, const int32_t *d_str_offsets, const int64_t *d_ids,
                            int64_t num_bins, int mask_mode, int use_strong, uint64_t key0, uint64_t key1,
                            int64_t *d_ids_ws, const int32_t *d_bag_offsets, int32_t bag_len, int64_t batch,
                            int world, int64_t *const *h_rows_dst, int32_t *const *h_begin_dst,
                            int32_t *const *h_end_dst, int max_ctas_per_sm, void *stream);

/* ---- the whole row-sharded forward step as ONE call (rf_sharded_bag_forward) -------------------- */
/* route -> barrier -> fused gather+pool into the SOURCE ranks' buffers over NVLink -> barrier -> combine, all stream  */
/* ordered, no host synchronisation, no torch / NCCL types: the exchange runs on peer-mapped memory the binder         */
/* provides (cudaMalloc + cudaIpc handles, cuMem fabric handles, an NCCL window, torch symmetric memory ...).          */
/* Every rank allocates rf_shard_exchange_bytes(...) bytes + a signal pad of >= 2 * 16 uint32 (zero-initialised once), */
/* maps every peer's two buffers, and fills the context with the pointers AS MAPPED INTO ITS OWN process.              */
/* `step` counts 1, 2, 3, ... identically on every rank (the barriers compare it; no reset between steps).             */
/* Keys: strings (d_bytes + d_str_offsets, hashed on the fly) or pre-hashed d_ids; bags: CSR d_bag_offsets or dense    */
/* bag_len.  d_shard: this rank's rows (id % world == rank, local row id / world), [shard_rows, dim] fp32.             */
/* Pads / semantics as rf_bag_forward; partial pools are summed in key order per owner and combined in rank order.     */
typedef struct rf_shard_ctx {
    int32_t rank, world;           /* world <= 16 (one NVSwitch box), one rank per GPU                         */
    int64_t max_batch;             /* bags per rank per step the exchange buffers were sized for               */
    int64_t max_keys;              /* keys per rank per step the exchange buffers were sized for               */
    int32_t dim;
    int32_t reserved;
    void *peer_exchange[16];       /* [world] every rank's exchange buffer, mapped into this process           */
    uint32_t *peer_signals[16];    /* [world] every rank's signal pad, mapped into this process                */
} rf_shard_ctx;
int64_t rf_shard_exchange_bytes(int world, int64_t max_batch, int64_t max_keys, int32_t dim);
int rf_sharded_bag_forward(const rf_shard_ctx *ctx, const uint8_t *d_bytes, const int32_t *d_str_offsets,
                           const int64_t *d_ids, int64_t num_bins, int mask_mode, int use_strong, uint64_t key0,
                           uint64_t key1, const int32_t *d_bag_offsets, int32_t bag_len, int64_t batch,
                           const float *d_shard, int64_t shard_rows, int combiner, uint64_t step, float *d_out,
                           int64_t out_stride, void *stream);

/* out[b] = reduce_{g<world, in rank order} partials[g][b][:]; avg divides by the bag's key count */
int rf_combine_partials(const float *d_partials, int world, int64_t batch, int32_t dim, int combiner,
                        int32_t bag_len, const int32_t *d_bag_offsets, float *d_out, int64_t out_stride,
                        void *stream);

/* ---- backward of a pooled bag, fused with the row update ("next" row) ------------ */
/* W[ids[k]] += alpha * grad_out[bag(k)] (x 1/count for avg) for every key k; alpha = -lr is the    */
/* SGD step of the reference's Embedding variables (pads included: row 0 collects their gradient). */
/* d_ids: the bucket ids the forward produced (rf_field_desc.ids_out, one table's slice).           */
int rf_bag_backward(const int64_t *d_ids, int64_t n_keys, const int32_t *d_bag_offsets, int32_t bag_len,
                    int64_t batch, const float *d_grad_out, int64_t grad_stride, int32_t dim, int combiner,
                    float alpha, float *d_table, void *stream);

/* Same gradient, fused with the Adam update the reference trains with: tf.keras.optimizers.Adam on   */
/* the Embedding variables (example/ranking_search/train.py:97-104).  Keras semantics: duplicate ids */
/* are summed first; then, with lazy == 0, EVERY row of the table decays its moments and moves       */
/* (m = m*b1 [+ g*(1-b1)], v = v*b2 [+ g*g*(1-b2)], w -= lr_t * m / (sqrt(v) + eps),                  */
/* lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)); lazy == 1 touches only the gathered rows (LazyAdam, not  */
/* the reference's semantics).  ids must be < table_rows <= 2^32 - 1; dim % 4 == 0; d_table, d_m, d_v */
/* are [table_rows, dim] fp32, 16-byte aligned.  Workspace: device memory of                           */
/* rf_bag_adam_workspace_bytes(n_keys, table_rows) bytes (-1 on bad arguments).                        */
typedef struct rf_adam_params {
    float lr, beta1, beta2, epsilon; /* Keras defaults: 1e-3, 0.9, 0.999, 1e-7 */
    int64_t step;                    /* t >= 1: iterations + 1                  */
    int32_t lazy;
    int32_t reserved;
    const float *d_lr_t;             /* NULL, or a device float holding lr * sqrt(1 - beta2^t) / (1 - beta1^t): read by the   */
                                     /* kernels instead of the value derived from `step` -- required when the call is being   */
                                     /* recorded into a CUDA graph (the host-side step would be frozen into the graph)        */
    uint32_t *d_live_rows;           /* NULL, or a caller-owned device bitmap of ceil(sum table_rows / 32) + 1 words, zero before */
                                     /* the first step and passed to EVERY step with the same table list (bit = running row   */
                                     /* number over the tables in call order): the rows that ever received a gradient.  A row */
                                     /* whose bit is clear still has m == v == 0, so the all-rows decay of non-lazy Adam       */
                                     /* skips it without reading it (bit-identical results; all ones is always safe).          */
} rf_adam_params;
/* One table of a multi-table update.  All tables of a call share `dim`; sum of table_rows <= 2^32 - 1, */
/* sum of n_keys <= 2^31 - 1.  grad_out: [batch, dim] view with row stride grad_stride (floats).         */
typedef struct rf_adam_field {
    const int64_t *ids;         /* [n_keys] row ids gathered from this table (rf_field_desc.ids_out slice) */
    const int32_t *bag_offsets; /* [batch + 1] jagged bags, or NULL for batch x bag_len                    */
    int64_t n_keys;
    int32_t bag_len;
    int32_t combiner;           /* RF_COMBINER_SUM | RF_COMBINER_AVG                                        */
    const float *grad_out;
    int64_t grad_stride;
    float *table, *m, *v;       /* [table_rows, dim] fp32                                                   */
    int64_t table_rows;
    int32_t dim;
    int32_t reserved;
} rf_adam_field;
/* ONE sort / select / update pass over all the tables (instead of one per table).                     */
int64_t rf_bag_adam_multi_workspace_bytes(const rf_adam_field *fields, int n_fields);
int rf_bag_backward_adam_multi(const rf_adam_field *fields, int n_fields, int64_t batch,
                               const rf_adam_params *params, void *d_workspace, int64_t workspace_bytes,
                               void *stream);
int64_t rf_bag_adam_workspace_bytes(int64_t n_keys, int64_t table_rows);
int rf_bag_backward_adam(const int64_t *d_ids, int64_t n_keys, const int32_t *d_bag_offsets, int32_t bag_len,
                         int64_t batch, const float *d_grad_out, int64_t grad_stride, int32_t dim, int combiner,
                         const rf_adam_params *params, float *d_table, float *d_m, float *d_v,
                         int64_t table_rows, void *d_workspace, int64_t workspace_bytes, void *stream);

/* min / max pooling backward (tf.reduce_min / tf.reduce_max over the bag, preprocess_layers.py:43-68; TensorFlow's       */
/* _MinOrMaxGrad): d_key_grads[k][d] = grad_out[bag(k)][d] * (table[ids[k]][d] == pooled[bag][d]) / (number of keys of the  */
/* bag with that equality).  One gradient row per key ([n_keys, dim], dense); apply it with rf_bag_backward /                */
/* rf_bag_backward_adam as sum pooling over bags of one key (bag_len = 1, batch = n_keys).  d_pooled: the forward's output.   */
int rf_bag_minmax_key_grads(const int64_t *d_ids, int64_t n_keys, const int32_t *d_bag_offsets, int32_t bag_len,
                            int64_t batch, const float *d_table, int32_t dim, const float *d_pooled,
                            int64_t pooled_stride, const float *d_grad_out, int64_t grad_stride, float *d_key_grads,
                            void *stream);

/* ---- backward of the dense contractions (CUDA-core fp32; what model.fit differentiates) -------- */
/* Gradient of scaled_dot_product_attention (layer_utils.py:4-24) w.r.t. q, k, v given d_grad_out =   */
/* dL/d(out); same layouts as rf_sdpa_forward.  A masked query row passes gradient to v only.        */
/* seq_len <= 64, head_dim <= 128, else RF_ERR_UNSUPPORTED.                                           */
int rf_sdpa_backward(const float *d_q, const float *d_k, const float *d_v, const float *d_mask,
                     const float *d_grad_out, int64_t n_batch_heads, int32_t seq_len, int32_t head_dim,
                     float *d_dq, float *d_dk, float *d_dv, void *stream);
/* The same with q, k, v read at a row pitch of row_pitch floats (column windows of one fused q|k|v      */
/* projection output) and dq, dk, dv written at grad_row_pitch; d_grad_out dense [n, seq_len, head_dim]. */
/* head_dim and the pitches multiples of 4, 16-byte aligned buffers.                                     */
int rf_sdpa_backward_strided(const float *d_q, const float *d_k, const float *d_v, int64_t row_pitch, const float *d_mask,
                             const float *d_grad_out, int64_t n_batch_heads, int32_t seq_len, int32_t head_dim,
                             float *d_dq, float *d_dk, float *d_dv, int64_t grad_row_pitch, void *stream);
/* The same on the tensor cores (TF32 operands, fp32 accumulate; warp-level mma.sync on operands held in shared memory, */
/* the softmax / delta arithmetic stays fp32): seq_len <= 64, head_dim in {32, 64, 96, 128}.                             */
int rf_sdpa_backward_tc(const float *d_q, const float *d_k, const float *d_v, int64_t row_pitch, const float *d_mask,
                        const float *d_grad_out, int64_t n_batch_heads, int32_t seq_len, int32_t head_dim,
                        float *d_dq, float *d_dk, float *d_dv, int64_t grad_row_pitch, void *stream);
/* Gradient of batch_neg_sample_scaled_multi_class_ce_loss (match_losses.py:150-165) w.r.t. query   */
/* and doc (either output may be NULL), times `upstream` (dL/d loss).  d_lse: the per-row            */
/* log-sum-exp rf_inbatch_rowstats[_tc] produced for the same inputs.  dim <= 512.  Deterministic.   */
int rf_inbatch_softmax_ce_backward(const float *d_query, const float *d_doc, const float *d_y,
                                   const float *d_lse, int64_t batch, int32_t dim, float scale,
                                   float upstream, float *d_grad_query, float *d_grad_doc, void *stream);

/* One [batch x batch] BLOCK of a wider logits matrix (data-parallel towers: every rank's queries against the docs  */
/* all-gathered from all ranks).  d_lse is the log-sum-exp over the WHOLE row (all blocks); the positives sit on the */
/* diagonal only in the rank's own block (positives_on_diagonal = 1), other blocks hold negatives only (0).          */
int rf_inbatch_softmax_ce_backward_block(const float *d_query, const float *d_doc, const float *d_y, const float *d_lse,
                                         int64_t batch, int32_t dim, float scale, float upstream,
                                         int positives_on_diagonal, float *d_grad_query, float *d_grad_doc,
                                         void *stream);

/* The same gradients on the tensor cores: the three contractions (S = Q D^T, dQ = C D, dD = C^T Q) run through          */
/* rf_dense_forward_tc (tcgen05, TF32 operands) over slabs of query rows; only a [slab x batch] piece of the coefficient */
/* matrix (<= 512 MiB, plus its transpose) exists at a time.  batch % 4 == 0, dim % 4 == 0; both gradients are produced.                      */
/* positives_on_diagonal as in rf_inbatch_softmax_ce_backward_block.                                                      */
int64_t rf_inbatch_ce_backward_tc_workspace_bytes(int64_t batch, int32_t dim);
int rf_inbatch_softmax_ce_backward_tc(const float *d_query, const float *d_doc, const float *d_y, const float *d_lse,
                                      int64_t batch, int32_t dim, float scale, float upstream, int positives_on_diagonal,
                                      void *d_workspace, int64_t workspace_bytes, float *d_grad_query, float *d_grad_doc,
                                      void *stream);

/* ---- training-time passes of one tower stage  y = act(BatchNormalization_batch(x) W + b) ------------------------------ */
/* (backend/blocks/mlp.py:4-15 under model.fit).  The stage's three GEMMs are rf_dense_forward_tc; these are the          */
/* HBM-bound column passes around them.  All buffers fp32 on the device; workspace: rf_tower_train_workspace_bytes(rows,   */
/* max(dim, units)) bytes.  Deterministic (ordered partial sums).                                                         */
int64_t rf_tower_train_workspace_bytes(int64_t rows, int32_t dim);
/* Batch mean and BIASED variance of every column of x [rows, dim] (row pitch ldx floats), Keras BatchNormalization with   */
/* training=True.  d_x_t: NULL, or [dim, rows] receiving x transposed in the same pass (operand of dW = X^T dZ).           */
int rf_column_stats(const float *d_x, int64_t rows, int32_t dim, int64_t ldx, float *d_mean, float *d_var, float *d_x_t,
                    void *d_workspace, int64_t workspace_bytes, void *stream);
/* dZ = dY * act'(z) with the derivative taken from the stage's OUTPUT y = act(z) (none, relu, selu, tanh, sigmoid);        */
/* d_grad_pre [rows, units] (may be NULL for activation none: dZ is dY), d_grad_pre_t NULL or [units, rows] (dZ           */
/* transposed), d_grad_bias [units] = column sums of dZ.                                                                  */
int rf_activation_backward(const float *d_grad_out, const float *d_out, int64_t rows, int32_t units, int activation,
                           float *d_grad_pre, float *d_grad_pre_t, float *d_grad_bias, void *d_workspace,
                           int64_t workspace_bytes, void *stream);
/* BatchNormalization backward on batch statistics: given dXhat [rows, dim] (gradient w.r.t. the normalised + affine      */
/* output), x, the batch mean, rstd = rsqrt(var + eps) and scale = gamma * rstd:                                           */
/*   dbeta = colsum(dXhat), dgamma = colsum(dXhat * xn), dX = scale * (dXhat - dbeta / rows - xn * dgamma / rows),          */
/*   xn = (x - mean) * rstd.  dim and ldx multiples of 4.                                                                  */
int rf_batchnorm_backward(const float *d_grad_normed, const float *d_x, int64_t ldx, const float *d_mean, const float *d_rstd,
                          const float *d_scale, int64_t rows, int32_t dim, float *d_grad_gamma, float *d_grad_beta,
                          float *d_grad_x, void *d_workspace, int64_t workspace_bytes, void *stream);

/* ---- vocabulary lookup / bucketisation ------------------------------------ */
/* Keras StringLookup / IntegerLookup(vocabulary=vocabs, output_mode="int") as LookupEmbedding builds */
/* them (backend/layers/preprocess_layers.py:148-150): term i -> i + 1, out-of-vocabulary -> 0.      */
/* The table is caller-owned device memory; exact (a hash match is confirmed against the term).     */
typedef struct rf_vocab_desc {
    const uint8_t *term_bytes;    /* string vocabulary: arena of the terms (+16 readable bytes), else NULL */
    const int32_t *term_offsets;  /* [n_terms + 1] byte offsets into term_bytes, else NULL                */
    const int64_t *term_ints;     /* integer vocabulary: [n_terms], else NULL                             */
    uint64_t *slots;              /* [capacity] open-addressing table, filled by rf_vocab_build           */
    int64_t capacity;             /* power of two, >= max(2, 2 * n_terms)                                 */
    int64_t n_terms;              /* terms must be distinct (the caller checks, as Keras does)            */
} rf_vocab_desc;
int rf_vocab_build(const rf_vocab_desc *vocab, void *stream);
int rf_vocab_lookup_strings(const rf_vocab_desc *vocab, const uint8_t *d_bytes, const int32_t *d_str_offsets,
                            int64_t n_items, int64_t *d_ids_out, void *stream);
int rf_vocab_lookup_int64(const rf_vocab_desc *vocab, const int64_t *d_values, int64_t n_items,
                          int64_t *d_ids_out, void *stream);
/* Keras Discretization(bin_boundaries) (preprocess_layers.py:187): id = #boundaries <= x            */
/* (upper_bound with `x < boundary`; NaN -> n_boundaries).  Boundaries ascending, device fp32.        */
int rf_bucketize_f32(const float *d_values, int64_t n_items, const float *d_boundaries, int32_t n_boundaries,
                     int64_t *d_ids_out, void *stream);

/* Cap the fused kernel's grid at ctas_per_sm x #SMs (0 = no cap; it then walks its tiles            */
/* grid-stride).  Leaves room on every SM for kernels of other streams (the sharded pipeline).     */
int rf_set_bag_grid_limit(int ctas_per_sm);

/* Number of kernels launched by this library since load (bench.py's gpu_launches counter).   */
int64_t rf_launch_count(void);

/* Host-side check of the division-by-invariant used for `h mod bins` (tests only). */
uint64_t rf_debug_fastmod(uint64_t x, uint64_t d);

#ifdef __cplusplus
}
#endif
#endif /* RF_B200_H_ */
