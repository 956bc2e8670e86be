/*
 * rf_oracle.c -- CPU ORACLE for the RecommendFlow feature-to-embedding hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (recommendflow_b200/) never imports, links or executes anything in oracle/.
 *
 * It restates, in plain C, the algorithm the reference runs through TensorFlow/Keras
 * library ops (TensorFlow itself is a third-party dependency that is NOT vendored under
 * /root/reference and is not installed here; no version is pinned by the reference --
 * inferred TF 2.6-2.8 from the Keras symbols imported at
 * backend/layers/preprocess_layers.py:11):
 *
 *   - Keras `Hashing(num_bins, mask_value, salt)`            preprocess_layers.py:89-90,95
 *       salt=None  -> tf.strings.to_hash_bucket_fast   = farmhashna::Hash64 (Fingerprint64) mod bins
 *       salt=[a,b] -> tf.strings.to_hash_bucket_strong = SipHash-2-4(key a,b) mod bins
 *   - Keras `Embedding` gather + combiner over axis 1       preprocess_layers.py:43-68
 *       (pads are NOT masked: id 0 -> row 0 is pooled in; avg divides by padded L)
 *   - DoubleHashingEmbedding concat                         preprocess_layers.py:94-97
 *   - scaled_dot_product_attention                          backend/layers/layer_utils.py:4-24
 *   - batch_neg_sample_scaled_multi_class_ce_loss           backend/lossess/match_losses.py:150-165
 *
 * PARITY PINNING: the reference has no tests, fixtures or golden vectors,
 * and TF cannot be run here.  The hash functions are pinned against PUBLIC TF/Keras
 * known-answer vectors (tests/test_oracle_kat.py); FarmHash inputs longer than 16 bytes
 * are covered only by two independently written restatements agreeing (this file and
 * oracle/pyhash.py) -- for those branches parity is UNPINNED.
 *
 * Build: make -C oracle   (gcc -O3 -fopenmp -shared -fPIC)
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------ */
/* FarmHash Fingerprint64 (farmhashna::Hash64), published algorithm, SURVEY.md App. B     */
/* ------------------------------------------------------------------------------------ */
static const uint64_t K0 = 0xc3a5c85c97cb3127ULL;
static const uint64_t K1 = 0xb492b66fbe98f273ULL;
static const uint64_t K2 = 0x9ae16a3b2f90404fULL;

static inline uint64_t f64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint64_t f32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t ror(uint64_t v, int s) { return s == 0 ? v : ((v >> s) | (v << (64 - s))); }
static inline uint64_t smix(uint64_t v) { return v ^ (v >> 47); }

static inline uint64_t hl16(uint64_t u, uint64_t v, uint64_t mul) {
    uint64_t a = (u ^ v) * mul;
    a ^= (a >> 47);
    uint64_t b = (v ^ a) * mul;
    b ^= (b >> 47);
    return b * mul;
}

static uint64_t farm_0_16(const uint8_t *s, size_t n) {
    if (n >= 8) {
        uint64_t mul = K2 + n * 2;
        uint64_t a = f64(s) + K2;
        uint64_t b = f64(s + n - 8);
        uint64_t c = ror(b, 37) * mul + a;
        uint64_t d = (ror(a, 25) + b) * mul;
        return hl16(c, d, mul);
    }
    if (n >= 4) {
        uint64_t mul = K2 + n * 2;
        uint64_t a = f32(s);
        return hl16(n + (a << 3), f32(s + n - 4), mul);
    }
    if (n > 0) {
        uint8_t a = s[0], b = s[n >> 1], c = s[n - 1];
        uint32_t y = (uint32_t)a + ((uint32_t)b << 8);
        uint32_t z = (uint32_t)n + ((uint32_t)c << 2);
        return smix(y * K2 ^ z * K0) * K2;
    }
    return K2;
}

static uint64_t farm_17_32(const uint8_t *s, size_t n) {
    uint64_t mul = K2 + n * 2;
    uint64_t a = f64(s) * K1;
    uint64_t b = f64(s + 8);
    uint64_t c = f64(s + n - 8) * mul;
    uint64_t d = f64(s + n - 16) * K2;
    return hl16(ror(a + b, 43) + ror(c, 30) + d, a + ror(b + K2, 18) + c, mul);
}

static uint64_t farm_33_64(const uint8_t *s, size_t n) {
    uint64_t mul = K2 + n * 2;
    uint64_t a = f64(s) * K2;
    uint64_t b = f64(s + 8);
    uint64_t c = f64(s + n - 8) * mul;
    uint64_t d = f64(s + n - 16) * K2;
    uint64_t y = ror(a + b, 43) + ror(c, 30) + d;
    uint64_t z = hl16(y, a + ror(b + K2, 18) + c, mul);
    uint64_t e = f64(s + 16) * mul;
    uint64_t f = f64(s + 24);
    uint64_t g = (y + f64(s + n - 32)) * mul;
    uint64_t h = (z + f64(s + n - 24)) * mul;
    return hl16(ror(e + f, 43) + ror(g, 30) + h, e + ror(f + a, 18) + g, mul);
}

typedef struct { uint64_t first, second; } pair64;

static inline pair64 weak32(const uint8_t *p, uint64_t a, uint64_t b) {
    uint64_t w = f64(p), x = f64(p + 8), y = f64(p + 16), z = f64(p + 24);
    a += w;
    b = ror(b + a + z, 21);
    uint64_t c = a;
    a += x;
    a += y;
    b += ror(a, 44);
    pair64 r = { a + z, b + c };
    return r;
}

uint64_t rfo_fingerprint64(const uint8_t *s, uint64_t n) {
    if (n <= 16) return farm_0_16(s, n);
    if (n <= 32) return farm_17_32(s, n);
    if (n <= 64) return farm_33_64(s, n);
    const uint64_t seed = 81;
    uint64_t x = seed;
    uint64_t y = seed * K1 + 113;
    uint64_t z = smix(y * K2 + 113) * K2;
    pair64 v = { 0, 0 }, w = { 0, 0 };
    x = x * K2 + f64(s);
    const uint8_t *end = s + ((n - 1) / 64) * 64;
    const uint8_t *last64 = end + ((n - 1) & 63) - 63;
    do {
        x = ror(x + y + v.first + f64(s + 8), 37) * K1;
        y = ror(y + v.second + f64(s + 48), 42) * K1;
        x ^= w.second;
        y += v.first + f64(s + 40);
        z = ror(z + w.first, 33) * K1;
        v = weak32(s, v.second * K1, x + w.first);
        w = weak32(s + 32, z + w.second, y + f64(s + 16));
        { uint64_t t = z; z = x; x = t; }
        s += 64;
    } while (s != end);
    uint64_t mul = K1 + ((z & 0xff) << 1);
    s = last64;
    w.first += ((n - 1) & 63);
    v.first += w.first;
    w.first += v.first;
    x = ror(x + y + v.first + f64(s + 8), 37) * mul;
    y = ror(y + v.second + f64(s + 48), 42) * mul;
    x ^= w.second * 9;
    y += v.first * 9 + f64(s + 40);
    z = ror(z + w.first, 33) * mul;
    v = weak32(s, v.second * mul, x + w.first);
    w = weak32(s + 32, z + w.second, y + f64(s + 16));
    { uint64_t t = z; z = x; x = t; }
    return hl16(hl16(v.first, w.first, mul) + smix(y) * K0 + z,
                hl16(v.second, w.second, mul) + x, mul);
}

/* ------------------------------------------------------------------------------------ */
/* SipHash-2-4 (highwayhash::SipHash as used by TF StrongKeyedHash), published algorithm  */
/* ------------------------------------------------------------------------------------ */
static inline uint64_t rol(uint64_t v, int s) { return (v << s) | (v >> (64 - s)); }

#define SIPROUND do {                                   \
        v0 += v1; v1 = rol(v1, 13); v1 ^= v0; v0 = rol(v0, 32); \
        v2 += v3; v3 = rol(v3, 16); v3 ^= v2;                   \
        v0 += v3; v3 = rol(v3, 21); v3 ^= v0;                   \
        v2 += v1; v1 = rol(v1, 17); v1 ^= v2; v2 = rol(v2, 32); \
    } while (0)

uint64_t rfo_siphash24(uint64_t k0, uint64_t k1, const uint8_t *s, uint64_t n) {
    uint64_t v0 = k0 ^ 0x736f6d6570736575ULL;
    uint64_t v1 = k1 ^ 0x646f72616e646f6dULL;
    uint64_t v2 = k0 ^ 0x6c7967656e657261ULL;
    uint64_t v3 = k1 ^ 0x7465646279746573ULL;
    const uint8_t *end = s + (n & ~(uint64_t)7);
    for (const uint8_t *p = s; p != end; p += 8) {
        uint64_t m = f64(p);
        v3 ^= m; SIPROUND; SIPROUND; v0 ^= m;
    }
    uint64_t m = (n & 0xff) << 56;
    for (uint64_t i = 0; i < (n & 7); ++i) m |= (uint64_t)end[i] << (8 * i);
    v3 ^= m; SIPROUND; SIPROUND; v0 ^= m;
    v2 ^= 0xff;
    SIPROUND; SIPROUND; SIPROUND; SIPROUND;
    return v0 ^ v1 ^ v2 ^ v3;
}

/* ------------------------------------------------------------------------------------ */
/* tf.as_string(int64): base 10, '-' for negatives, no padding                            */
/* ------------------------------------------------------------------------------------ */
int rfo_as_string_i64(int64_t v, uint8_t *out /* >= 20 bytes */) {
    char tmp[24];
    int n = snprintf(tmp, sizeof tmp, "%lld", (long long)v);
    memcpy(out, tmp, (size_t)n);
    return n;
}

/* ------------------------------------------------------------------------------------ */
/* Keras Hashing._hash_values_to_bins                                                     */
/*   bins' = num_bins-1 if (mask given and num_bins>1) else num_bins                      */
/*   h     = strong ? siphash24(k0,k1,s) : fingerprint64(s);  id = h mod bins'            */
/*   masked:  id = (s == mask_value) ? 0 : id + 1                                         */
/* ------------------------------------------------------------------------------------ */
static inline int64_t keras_bucket(const uint8_t *s, uint64_t n, int64_t num_bins, int use_strong,
                                   uint64_t k0, uint64_t k1, int has_mask, int is_mask) {
    uint64_t bins = (uint64_t)num_bins;
    int masking = has_mask && num_bins > 1;
    if (masking) bins -= 1;
    uint64_t h = use_strong ? rfo_siphash24(k0, k1, s, n) : rfo_fingerprint64(s, n);
    int64_t id = (int64_t)(h % bins);
    if (masking) id = is_mask ? 0 : id + 1;
    return id;
}

/* strings: bytes arena + int32 offsets[n+1]; mask_value = (mask, mask_len) when has_mask */
void rfo_hash_strings(const uint8_t *bytes, const int32_t *offs, int64_t n, int64_t num_bins,
                      int use_strong, uint64_t k0, uint64_t k1,
                      int has_mask, const uint8_t *mask, int32_t mask_len, int64_t *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t *s = bytes + offs[i];
        uint64_t len = (uint64_t)(offs[i + 1] - offs[i]);
        int is_mask = has_mask && (int64_t)len == mask_len && (len == 0 || memcmp(s, mask, len) == 0);
        out[i] = keras_bucket(s, len, num_bins, use_strong, k0, k1, has_mask, is_mask);
    }
}

/* int64 inputs: mask compares the integer, then as_string, then hash */
void rfo_hash_ints(const int64_t *vals, int64_t n, int64_t num_bins, int use_strong,
                   uint64_t k0, uint64_t k1, int has_mask, int64_t mask_value, int64_t *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint8_t buf[24];
        int len = rfo_as_string_i64(vals[i], buf);
        int is_mask = has_mask && vals[i] == mask_value;
        out[i] = keras_bucket(buf, (uint64_t)len, num_bins, use_strong, k0, k1, has_mask, is_mask);
    }
}

/* ------------------------------------------------------------------------------------ */
/* EmbeddingBag (preprocess_layers.py:43-68): E = W[ids]; reduce over axis 1.             */
/* combiner: 0 sum, 1 avg, 2 min, 3 max.  Accumulation is sequential in index order       */
/* (l = 0..L-1) in fp32; avg = sum / L with an fp32 division.                             */
/* Bags: dense [B,L] when bag_offs == NULL (every position counts, pads included),        */
/* else CSR bag_offs[B+1] over the flat id list (jagged mode; empty bag -> 0).            */
/* ------------------------------------------------------------------------------------ */
void rfo_bag_pool(const int64_t *ids, int64_t B, int64_t L, const int32_t *bag_offs,
                  const float *W, int64_t N, int64_t D, int combiner,
                  float *out, int64_t out_stride) {
    (void)N;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b) {
        int64_t lo = bag_offs ? bag_offs[b] : b * L;
        int64_t hi = bag_offs ? bag_offs[b + 1] : (b + 1) * L;
        float *o = out + b * out_stride;
        if (hi <= lo) { for (int64_t d = 0; d < D; ++d) o[d] = 0.0f; continue; }
        for (int64_t d = 0; d < D; ++d) {
            const float first = W[ids[lo] * D + d];
            float acc = (combiner <= 1) ? 0.0f + first : first;
            for (int64_t i = lo + 1; i < hi; ++i) {
                float x = W[ids[i] * D + d];
                if (combiner <= 1) acc += x;
                else if (combiner == 2) acc = x < acc ? x : acc;
                else acc = x > acc ? x : acc;
            }
            if (combiner == 1) acc = acc / (float)(hi - lo);
            o[d] = acc;
        }
    }
}

/* combiner "null": plain gather, out[i, :] = W[ids[i], :] */
void rfo_gather_rows(const int64_t *ids, int64_t n, const float *W, int64_t D, float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) memcpy(out + i * D, W + ids[i] * D, (size_t)D * 4);
}

/* Fused reference forward of one hashed field with T tables (T=2: DoubleHashingEmbedding,
 * preprocess_layers.py:94-97).  out[b, t*D:(t+1)*D] = pool_l W_t[hash_t(x[b,l])].        */
void rfo_hashed_bag_forward(const uint8_t *bytes, const int32_t *offs, int64_t B, int64_t L,
                            const int32_t *bag_offs, int T, const float *const *W,
                            const int64_t *num_bins, const int *use_strong,
                            const uint64_t *k0, const uint64_t *k1, int has_mask,
                            int64_t D, int combiner, float *out, int64_t out_stride) {
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b) {
        int64_t lo = bag_offs ? bag_offs[b] : b * L;
        int64_t hi = bag_offs ? bag_offs[b + 1] : (b + 1) * L;
        for (int t = 0; t < T; ++t) {
            float *o = out + b * out_stride + (int64_t)t * D;
            for (int64_t d = 0; d < D; ++d) o[d] = 0.0f;
            for (int64_t i = lo; i < hi; ++i) {
                const uint8_t *s = bytes + offs[i];
                uint64_t len = (uint64_t)(offs[i + 1] - offs[i]);
                int64_t id = keras_bucket(s, len, num_bins[t], use_strong[t], k0[t], k1[t],
                                          has_mask, has_mask && len == 0);
                const float *row = W[t] + id * D;
                if (combiner <= 1 || i == lo) {
                    if (i == lo && combiner > 1) for (int64_t d = 0; d < D; ++d) o[d] = row[d];
                    else for (int64_t d = 0; d < D; ++d) o[d] += row[d];
                } else if (combiner == 2) {
                    for (int64_t d = 0; d < D; ++d) o[d] = row[d] < o[d] ? row[d] : o[d];
                } else {
                    for (int64_t d = 0; d < D; ++d) o[d] = row[d] > o[d] ? row[d] : o[d];
                }
            }
            if (combiner == 1 && hi > lo) {
                float cnt = (float)(hi - lo);
                for (int64_t d = 0; d < D; ++d) o[d] = o[d] / cnt;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* scaled_dot_product_attention (layer_utils.py:4-24), q,k,v [NB,S,dh] fp32,              */
/* mask [NB,S] (the reference's [...,S,1] mask broadcasts over KEYS, i.e. it masks whole  */
/* QUERY rows: a masked row gets logits -4294967295 everywhere -> uniform attention).     */
/* Accumulation in double so the oracle is the tight side of the tolerance.               */
/* ------------------------------------------------------------------------------------ */
void rfo_sdpa(const float *q, const float *k, const float *v, const float *mask,
              int64_t NB, int64_t S, int64_t dh, float *out) {
    const float scale_den = sqrtf((float)dh);
#pragma omp parallel for schedule(static)
    for (int64_t nb = 0; nb < NB; ++nb) {
        double *p = (double *)malloc(sizeof(double) * (size_t)S);
        for (int64_t i = 0; i < S; ++i) {
            const float *qi = q + (nb * S + i) * dh;
            int masked = mask && mask[nb * S + i] == 0.0f;
            double mx = -INFINITY;
            for (int64_t j = 0; j < S; ++j) {
                double acc = 0.0;
                const float *kj = k + (nb * S + j) * dh;
                for (int64_t d = 0; d < dh; ++d) acc += (double)qi[d] * (double)kj[d];
                double logit = (double)((float)acc / scale_den);
                if (masked) logit = (double)(-4294967295.0f);
                p[j] = logit;
                if (logit > mx) mx = logit;
            }
            double den = 0.0;
            for (int64_t j = 0; j < S; ++j) { p[j] = exp(p[j] - mx); den += p[j]; }
            for (int64_t d = 0; d < dh; ++d) {
                double acc = 0.0;
                for (int64_t j = 0; j < S; ++j) acc += p[j] * (double)v[(nb * S + j) * dh + d];
                out[(nb * S + i) * dh + d] = (float)(acc / den);
            }
        }
        free(p);
    }
}

/* ------------------------------------------------------------------------------------ */
/* batch_neg_sample_scaled_multi_class_ce_loss (match_losses.py:150-165)                  */
/*   S = q d^T; loss = mean_i( -log(exp(s*S_ii) / sum_j exp(s*S_ij)) * y_i )              */
/* no max-subtraction in the reference; here evaluated in double via log-sum-exp, which   */
/* is the same real number wherever the reference does not overflow.                      */
/* row_lse[i] = log(sum_j exp(s*S_ij)), diag[i] = S_ii are also returned.                 */
/* ------------------------------------------------------------------------------------ */
double rfo_inbatch_softmax_ce(const float *q, const float *d, const float *y, int64_t B, int64_t Dt,
                              double scale, double *row_lse, double *diag) {
    double total = 0.0;
#pragma omp parallel for schedule(static) reduction(+:total)
    for (int64_t i = 0; i < B; ++i) {
        const float *qi = q + i * Dt;
        double mx = -INFINITY, sii = 0.0;
        double *row = (double *)malloc(sizeof(double) * (size_t)B);
        for (int64_t j = 0; j < B; ++j) {
            const float *dj = d + j * Dt;
            double acc = 0.0;
            for (int64_t t = 0; t < Dt; ++t) acc += (double)qi[t] * (double)dj[t];
            row[j] = scale * acc;
            if (row[j] > mx) mx = row[j];
            if (j == i) sii = acc;
        }
        double den = 0.0;
        for (int64_t j = 0; j < B; ++j) den += exp(row[j] - mx);
        double lse = mx + log(den);
        free(row);
        if (row_lse) row_lse[i] = lse;
        if (diag) diag[i] = sii;
        total += -(scale * sii - lse) * (double)y[i];
    }
    return total / (double)B;
}

/* ------------------------------------------------------------------------------------ */
/* Backward of the pooled bag + tf.keras.optimizers.Adam on the Embedding variable            */
/* (example/ranking_search/train.py:97-104; Keras OptimizerV2 Adam._resource_apply_sparse     */
/* after _deduplicate_indexed_slices): duplicate rows' gradients are summed (here in key      */
/* order), then EVERY row decays its moments and moves; touched rows add the gradient terms.  */
/* lazy != 0: touched rows only.  G: caller-zeroed scratch [rows, D]; touched: zeroed [rows]. */
/* ------------------------------------------------------------------------------------ */
void rfo_bag_backward_adam(const int64_t *ids, int64_t n_keys, const int32_t *bag_offs, int64_t bag_len,
                           int64_t B, const float *grad, int64_t D, int avg, float lr, float beta1,
                           float beta2, float eps, int64_t step, int lazy, float *W, float *M, float *V,
                           int64_t rows, float *G, uint8_t *touched) {
    int64_t b = 0;
    for (int64_t k = 0; k < n_keys; ++k) {
        float scale = 1.0f;
        if (bag_offs) {
            while (b + 1 < B && bag_offs[b + 1] <= k) ++b;
            if (avg) scale = 1.0f / (float)(bag_offs[b + 1] - bag_offs[b]);
        } else {
            b = k / bag_len;
            if (avg) scale = 1.0f / (float)bag_len;
        }
        float *g = G + ids[k] * D;
        const float *src = grad + b * D;
        if (!touched[ids[k]]) {
            touched[ids[k]] = 1;
            for (int64_t j = 0; j < D; ++j) g[j] = avg ? src[j] * scale : src[j];
        } else {
            for (int64_t j = 0; j < D; ++j) g[j] = g[j] + (avg ? src[j] * scale : src[j]);
        }
    }
    const float b1p = powf(beta1, (float)step), b2p = powf(beta2, (float)step);
    const float lr_t = lr * sqrtf(1.0f - b2p) / (1.0f - b1p);
    const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
    for (int64_t r = 0; r < rows; ++r) {
        if (lazy && !touched[r]) continue;
        for (int64_t j = 0; j < D; ++j) {
            const int64_t at = r * D + j;
            float m = M[at] * beta1, v = V[at] * beta2;
            if (touched[r]) {
                const float g = G[at];
                m = m + g * omb1;
                v = v + (g * g) * omb2;
            }
            M[at] = m;
            V[at] = v;
            W[at] = W[at] - (lr_t * m) / (sqrtf(v) + eps);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Keras Dense: out[M,N] = act(x[M,K] . W[K,N] + b[N]), W = the Keras kernel [in, units]        */
/* (backend/blocks/mlp.py:4-15: keras.layers.Dense(units, activation); attention_layers.py:141). */
/* fp32 accumulation in k order (what TF's CPU kernel class does), one output row at a time.     */
/* act: 0 none, 1 relu, 2 selu, 3 tanh, 4 sigmoid, 5 gelu (erf form, the Keras default).         */
/* ------------------------------------------------------------------------------------------ */
void rfo_dense(const float *x, int64_t M, int64_t K, const float *W, const float *b, int64_t N, int act, float *out) {
#pragma omp parallel
    {
        float *acc = (float *)malloc((size_t)N * sizeof(float));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < M; ++i) {
            for (int64_t j = 0; j < N; ++j) acc[j] = 0.0f;
            const float *xi = x + i * K;
            for (int64_t k = 0; k < K; ++k) {
                const float a = xi[k];
                const float *w = W + k * N;
                for (int64_t j = 0; j < N; ++j) acc[j] += a * w[j];
            }
            float *o = out + i * N;
            for (int64_t j = 0; j < N; ++j) {
                float z = acc[j] + (b ? b[j] : 0.0f);
                switch (act) {
                    case 1: z = z > 0.0f ? z : 0.0f; break;
                    case 2: z = z > 0.0f ? 1.0507009873554805f * z : 1.0507009873554805f * 1.6732632423543772f * expm1f(z); break;
                    case 3: z = tanhf(z); break;
                    case 4: z = 1.0f / (1.0f + expf(-z)); break;
                    case 5: z = 0.5f * z * (1.0f + erff(z * 0.70710678118654752f)); break;
                    default: break;
                }
                o[j] = z;
            }
        }
        free(acc);
    }
}

int rfo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void rfo_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
