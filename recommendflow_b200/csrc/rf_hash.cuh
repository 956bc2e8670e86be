// rf_hash.cuh -- device-side FarmHash Fingerprint64, SipHash-2-4, tf.as_string(int64) and the
// Keras `Hashing` bucket rule, for sm_100a.
//
// What it must equal, bit for bit: Keras `Hashing(num_bins, mask_value, salt)` as configured
// at /root/reference/backend/layers/preprocess_layers.py:89-90 --
//   salt=None  -> tf.strings.to_hash_bucket_fast   (farmhashna::Hash64 mod bins)
//   salt=[a,b] -> tf.strings.to_hash_bucket_strong (SipHash-2-4 keyed (a,b) mod bins)
// Written from the published algorithms; structured around a
// "byte source" that serves unaligned little-endian fetches out of 4-byte-aligned words, so the
// same code hashes keys staged in shared memory and keys read straight from global memory.
#pragma once
#include <stdint.h>

namespace rf {

// ------------------------------------------------------------------------------------------
// h mod d for a launch-invariant d: multiply-high by a precomputed 65-bit reciprocal
// (round-up method, "add" form), no 64-bit hardware division on the device.
// ------------------------------------------------------------------------------------------
struct FastMod {
    uint64_t d;      // divisor (>= 1)
    uint64_t magic;  // low 64 bits of the 65-bit reciprocal, 0 for powers of two
    uint32_t shift;  // post-shift
    uint32_t is_one; // d == 1
};

#if defined(__CUDACC__)
#define RF_HD __host__ __device__ __forceinline__
#else
#define RF_HD inline
#endif

inline FastMod make_fastmod(uint64_t d) {
    FastMod m;
    m.d = d;
    m.is_one = (d == 1);
    m.magic = 0;
    m.shift = 0;
    if (d <= 1) return m;
    const uint32_t fl = 63u - (uint32_t)__builtin_clzll(d);
    if ((d & (d - 1)) == 0) {  // power of two: t = x >> 1, then >> (fl - 1)
        m.shift = fl - 1;
        return m;
    }
    const unsigned __int128 num = (unsigned __int128)1 << (64 + fl);
    uint64_t q = (uint64_t)(num / d);
    const uint64_t rem = (uint64_t)(num % d);
    q += q;
    const uint64_t twice = rem + rem;
    if (twice >= d || twice < rem) q += 1;
    m.magic = q + 1;
    m.shift = fl;
    return m;
}

RF_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

RF_HD uint64_t fastmod(uint64_t x, const FastMod &m) {
    if (m.is_one) return 0;
    const uint64_t q0 = mulhi64(m.magic, x);
    const uint64_t t = ((x - q0) >> 1) + q0;
    const uint64_t q = t >> m.shift;
    return x - q * m.d;
}

// Everything Keras `Hashing` needs to turn one key into a bucket id.
struct HashSpec {
    uint64_t k0, k1;      // SipHash key (use_strong)
    FastMod mod;          // divisor = num_bins, or num_bins - 1 when bucket 0 is reserved
    uint32_t use_strong;  // 1: SipHash-2-4, 0: FarmHash Fingerprint64
    uint32_t masking;     // 1: bucket 0 is reserved for the mask value (mask given and num_bins > 1)
};

inline HashSpec make_hash_spec(int64_t num_bins, bool has_mask, int use_strong, uint64_t k0, uint64_t k1) {
    HashSpec h;
    h.k0 = k0;
    h.k1 = k1;
    h.use_strong = use_strong ? 1u : 0u;
    h.masking = (has_mask && num_bins > 1) ? 1u : 0u;
    h.mod = make_fastmod((uint64_t)num_bins - (h.masking ? 1u : 0u));
    return h;
}

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------------
// Byte sources: `words` is 4-byte aligned, the key starts `off` bytes into it.  Fetches read
// whole aligned words and funnel-shift, so they may touch up to 7 bytes past the key's end:
// every buffer a source points into carries that much readable slack.
// ------------------------------------------------------------------------------------------
struct WordSrcShared {
    const uint32_t *words;
    uint32_t off;
    __device__ __forceinline__ uint32_t w(uint32_t i) const { return words[i]; }
};
struct WordSrcGlobal {
    const uint32_t *words;
    uint32_t off;
    __device__ __forceinline__ uint32_t w(uint32_t i) const { return __ldg(words + i); }
};

template <class Src>
__device__ __forceinline__ uint32_t fetch32(const Src &s, uint32_t pos) {
    const uint32_t p = s.off + pos;
    const uint32_t i = p >> 2, sh = (p & 3u) * 8u;
    return __funnelshift_r(s.w(i), s.w(i + 1), sh);
}
template <class Src>
__device__ __forceinline__ uint64_t fetch64(const Src &s, uint32_t pos) {
    const uint32_t p = s.off + pos;
    const uint32_t i = p >> 2, sh = (p & 3u) * 8u;
    const uint32_t a = s.w(i), b = s.w(i + 1), c = s.w(i + 2);
    return ((uint64_t)__funnelshift_r(b, c, sh) << 32) | __funnelshift_r(a, b, sh);
}
template <class Src>
__device__ __forceinline__ uint32_t fetch8(const Src &s, uint32_t pos) {
    const uint32_t p = s.off + pos;
    return (s.w(p >> 2) >> ((p & 3u) * 8u)) & 0xffu;
}

__device__ __forceinline__ uint64_t rotr64(uint64_t v, int s) { return (v >> s) | (v << (64 - s)); }
__device__ __forceinline__ uint64_t rotl64(uint64_t v, int s) { return (v << s) | (v >> (64 - s)); }

// ------------------------------------------------------------------------------------------
// FarmHash Fingerprint64 == farmhashna::Hash64
// ------------------------------------------------------------------------------------------
constexpr uint64_t kF0 = 0xc3a5c85c97cb3127ULL;
constexpr uint64_t kF1 = 0xb492b66fbe98f273ULL;
constexpr uint64_t kF2 = 0x9ae16a3b2f90404fULL;

__device__ __forceinline__ uint64_t shift_mix(uint64_t v) { return v ^ (v >> 47); }

__device__ __forceinline__ uint64_t hash_len16(uint64_t u, uint64_t v, uint64_t mul) {
    uint64_t a = (u ^ v) * mul;
    a ^= (a >> 47);
    uint64_t b = (v ^ a) * mul;
    b ^= (b >> 47);
    return b * mul;
}

struct U128 {
    uint64_t lo, hi;
};

template <class Src>
__device__ __forceinline__ U128 weak_hash32(const Src &s, uint32_t pos, uint64_t a, uint64_t b) {
    const uint64_t w = fetch64(s, pos), x = fetch64(s, pos + 8), y = fetch64(s, pos + 16),
                   z = fetch64(s, pos + 24);
    a += w;
    b = rotr64(b + a + z, 21);
    const uint64_t c = a;
    a += x;
    a += y;
    b += rotr64(a, 44);
    return U128{a + z, b + c};
}

template <class Src>
__device__ __noinline__ uint64_t farm_long(const Src &s, uint32_t n) {
    // n > 16: the three longer branches, kept out of line so that the common short-key path
    // stays small in registers and instruction footprint.
    if (n <= 32) {
        const uint64_t mul = kF2 + (uint64_t)n * 2;
        const uint64_t a = fetch64(s, 0) * kF1;
        const uint64_t b = fetch64(s, 8);
        const uint64_t c = fetch64(s, n - 8) * mul;
        const uint64_t d = fetch64(s, n - 16) * kF2;
        return hash_len16(rotr64(a + b, 43) + rotr64(c, 30) + d, a + rotr64(b + kF2, 18) + c, mul);
    }
    if (n <= 64) {
        const uint64_t mul = kF2 + (uint64_t)n * 2;
        const uint64_t a = fetch64(s, 0) * kF2;
        const uint64_t b = fetch64(s, 8);
        const uint64_t c = fetch64(s, n - 8) * mul;
        const uint64_t d = fetch64(s, n - 16) * kF2;
        const uint64_t y = rotr64(a + b, 43) + rotr64(c, 30) + d;
        const uint64_t z = hash_len16(y, a + rotr64(b + kF2, 18) + c, mul);
        const uint64_t e = fetch64(s, 16) * mul;
        const uint64_t f = fetch64(s, 24);
        const uint64_t g = (y + fetch64(s, n - 32)) * mul;
        const uint64_t h = (z + fetch64(s, n - 24)) * mul;
        return hash_len16(rotr64(e + f, 43) + rotr64(g, 30) + h, e + rotr64(f + a, 18) + g, mul);
    }
    const uint64_t seed = 81;
    uint64_t x = seed;
    uint64_t y = seed * kF1 + 113;
    uint64_t z = shift_mix(y * kF2 + 113) * kF2;
    U128 v{0, 0}, w{0, 0};
    x = x * kF2 + fetch64(s, 0);
    const uint32_t end = ((n - 1) / 64) * 64;
    const uint32_t last64 = end + ((n - 1) & 63) - 63;
    uint32_t p = 0;
    do {
        x = rotr64(x + y + v.lo + fetch64(s, p + 8), 37) * kF1;
        y = rotr64(y + v.hi + fetch64(s, p + 48), 42) * kF1;
        x ^= w.hi;
        y += v.lo + fetch64(s, p + 40);
        z = rotr64(z + w.lo, 33) * kF1;
        v = weak_hash32(s, p, v.hi * kF1, x + w.lo);
        w = weak_hash32(s, p + 32, z + w.hi, y + fetch64(s, p + 16));
        const uint64_t t = z;
        z = x;
        x = t;
        p += 64;
    } while (p != end);
    const uint64_t mul = kF1 + ((z & 0xff) << 1);
    p = last64;
    w.lo += ((n - 1) & 63);
    v.lo += w.lo;
    w.lo += v.lo;
    x = rotr64(x + y + v.lo + fetch64(s, p + 8), 37) * mul;
    y = rotr64(y + v.hi + fetch64(s, p + 48), 42) * mul;
    x ^= w.hi * 9;
    y += v.lo * 9 + fetch64(s, p + 40);
    z = rotr64(z + w.lo, 33) * mul;
    v = weak_hash32(s, p, v.hi * mul, x + w.lo);
    w = weak_hash32(s, p + 32, z + w.hi, y + fetch64(s, p + 16));
    const uint64_t t = z;
    z = x;
    x = t;
    return hash_len16(hash_len16(v.lo, w.lo, mul) + shift_mix(y) * kF0 + z,
                      hash_len16(v.hi, w.hi, mul) + x, mul);
}

template <class Src>
__device__ __forceinline__ uint64_t fingerprint64(const Src &s, uint32_t n) {
    if (n > 16) return farm_long(s, n);
    if (n >= 8) {
        const uint64_t mul = kF2 + (uint64_t)n * 2;
        const uint64_t a = fetch64(s, 0) + kF2;
        const uint64_t b = fetch64(s, n - 8);
        const uint64_t c = rotr64(b, 37) * mul + a;
        const uint64_t d = (rotr64(a, 25) + b) * mul;
        return hash_len16(c, d, mul);
    }
    if (n >= 4) {
        const uint64_t mul = kF2 + (uint64_t)n * 2;
        const uint64_t a = fetch32(s, 0);
        return hash_len16((uint64_t)n + (a << 3), fetch32(s, n - 4), mul);
    }
    if (n > 0) {
        const uint32_t a = fetch8(s, 0), b = fetch8(s, n >> 1), c = fetch8(s, n - 1);
        const uint32_t y = a + (b << 8);
        const uint32_t z = n + (c << 2);
        return shift_mix((uint64_t)y * kF2 ^ (uint64_t)z * kF0) * kF2;
    }
    return kF2;
}

// ------------------------------------------------------------------------------------------
// SipHash-2-4 (128-bit key as two little-endian uint64), as TF's StrongKeyedHash
// ------------------------------------------------------------------------------------------
#define RF_SIPROUND()            \
    do {                         \
        v0 += v1;                \
        v1 = rotl64(v1, 13);     \
        v1 ^= v0;                \
        v0 = rotl64(v0, 32);     \
        v2 += v3;                \
        v3 = rotl64(v3, 16);     \
        v3 ^= v2;                \
        v0 += v3;                \
        v3 = rotl64(v3, 21);     \
        v3 ^= v0;                \
        v2 += v1;                \
        v1 = rotl64(v1, 17);     \
        v1 ^= v2;                \
        v2 = rotl64(v2, 32);     \
    } while (0)

template <class Src>
__device__ __forceinline__ uint64_t siphash24(const Src &s, uint32_t n, uint64_t k0, uint64_t k1) {
    uint64_t v0 = k0 ^ 0x736f6d6570736575ULL;
    uint64_t v1 = k1 ^ 0x646f72616e646f6dULL;
    uint64_t v2 = k0 ^ 0x6c7967656e657261ULL;
    uint64_t v3 = k1 ^ 0x7465646279746573ULL;
    const uint32_t full = n & ~7u;
    for (uint32_t p = 0; p < full; p += 8) {
        const uint64_t m = fetch64(s, p);
        v3 ^= m;
        RF_SIPROUND();
        RF_SIPROUND();
        v0 ^= m;
    }
    const uint32_t rem = n & 7u;
    uint64_t m = (uint64_t)(n & 0xffu) << 56;
    if (rem) m |= fetch64(s, full) & ((1ULL << (rem * 8)) - 1ULL);
    v3 ^= m;
    RF_SIPROUND();
    RF_SIPROUND();
    v0 ^= m;
    v2 ^= 0xff;
    RF_SIPROUND();
    RF_SIPROUND();
    RF_SIPROUND();
    RF_SIPROUND();
    return v0 ^ v1 ^ v2 ^ v3;
}

// Two SipHash-2-4 of the SAME message under two keys, in lockstep: DoubleHashingEmbedding hashes every key twice
// (seeds[0], seeds[1]); one SipHash is a single dependent chain of 64-bit add / rotate / xor, so interleaving the two
// states doubles the instruction-level parallelism (C3: 228 features x 2 SipHash tables, D = 8 -- that launch is bound
// by hash issue latency, not by HBM).  Same arithmetic as siphash24, message words fetched once.
#define RF_SIPROUND2()                                                 \
    do {                                                               \
        a0 += a1; b0 += b1;                                            \
        a1 = rotl64(a1, 13); b1 = rotl64(b1, 13);                      \
        a1 ^= a0; b1 ^= b0;                                            \
        a0 = rotl64(a0, 32); b0 = rotl64(b0, 32);                      \
        a2 += a3; b2 += b3;                                            \
        a3 = rotl64(a3, 16); b3 = rotl64(b3, 16);                      \
        a3 ^= a2; b3 ^= b2;                                            \
        a0 += a3; b0 += b3;                                            \
        a3 = rotl64(a3, 21); b3 = rotl64(b3, 21);                      \
        a3 ^= a0; b3 ^= b0;                                            \
        a2 += a1; b2 += b1;                                            \
        a1 = rotl64(a1, 17); b1 = rotl64(b1, 17);                      \
        a1 ^= a2; b1 ^= b2;                                            \
        a2 = rotl64(a2, 32); b2 = rotl64(b2, 32);                      \
    } while (0)

template <class Src>
__device__ __forceinline__ void siphash24_x2(const Src &s, uint32_t n, uint64_t ka0, uint64_t ka1, uint64_t kb0, uint64_t kb1,
                                             uint64_t &ha, uint64_t &hb) {
    uint64_t a0 = ka0 ^ 0x736f6d6570736575ULL, a1 = ka1 ^ 0x646f72616e646f6dULL;
    uint64_t a2 = ka0 ^ 0x6c7967656e657261ULL, a3 = ka1 ^ 0x7465646279746573ULL;
    uint64_t b0 = kb0 ^ 0x736f6d6570736575ULL, b1 = kb1 ^ 0x646f72616e646f6dULL;
    uint64_t b2 = kb0 ^ 0x6c7967656e657261ULL, b3 = kb1 ^ 0x7465646279746573ULL;
    const uint32_t full = n & ~7u;
    for (uint32_t p = 0; p < full; p += 8) {
        const uint64_t m = fetch64(s, p);
        a3 ^= m; b3 ^= m;
        RF_SIPROUND2();
        RF_SIPROUND2();
        a0 ^= m; b0 ^= m;
    }
    const uint32_t rem = n & 7u;
    uint64_t m = (uint64_t)(n & 0xffu) << 56;
    if (rem) m |= fetch64(s, full) & ((1ULL << (rem * 8)) - 1ULL);
    a3 ^= m; b3 ^= m;
    RF_SIPROUND2();
    RF_SIPROUND2();
    a0 ^= m; b0 ^= m;
    a2 ^= 0xff; b2 ^= 0xff;
    RF_SIPROUND2();
    RF_SIPROUND2();
    RF_SIPROUND2();
    RF_SIPROUND2();
    ha = a0 ^ a1 ^ a2 ^ a3;
    hb = b0 ^ b1 ^ b2 ^ b3;
}

// Keras Hashing._hash_values_to_bins for one key under two strong hashes at once
template <class Src>
__device__ __forceinline__ void bucket_of_x2(const Src &src, uint32_t len, const HashSpec &ha, const HashSpec &hb, bool is_mask,
                                             uint32_t &ida, uint32_t &idb) {
    uint64_t xa, xb;
    siphash24_x2(src, len, ha.k0, ha.k1, hb.k0, hb.k1, xa, xb);
    ida = (uint32_t)fastmod(xa, ha.mod);
    idb = (uint32_t)fastmod(xb, hb.mod);
    if (ha.masking) ida = is_mask ? 0u : ida + 1u;
    if (hb.masking) idb = is_mask ? 0u : idb + 1u;
}

// Keras Hashing._hash_values_to_bins for one key
template <class Src>
__device__ __forceinline__ uint32_t bucket_of(const Src &src, uint32_t len, const HashSpec &h, bool is_mask) {
    const uint64_t x = h.use_strong ? siphash24(src, len, h.k0, h.k1) : fingerprint64(src, len);
    uint32_t id = (uint32_t)fastmod(x, h.mod);
    if (h.masking) id = is_mask ? 0u : id + 1u;
    return id;
}

// ------------------------------------------------------------------------------------------
// tf.as_string(int64): base-10 digits, '-' for negatives.  Writes into a 24-byte scratch of
// 4-byte words (6 words) and returns the length.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t format_int64(int64_t value, uint32_t *scratch6) {
    uint64_t mag = value < 0 ? (uint64_t)0 - (uint64_t)value : (uint64_t)value;
    uint8_t tmp[20];
    uint32_t nd = 0;
    do {
        const uint64_t q = mag / 10;
        tmp[nd++] = (uint8_t)('0' + (uint32_t)(mag - q * 10));
        mag = q;
    } while (mag != 0);
    const uint32_t neg = value < 0 ? 1u : 0u;
    const uint32_t len = nd + neg;
    uint8_t *dst = reinterpret_cast<uint8_t *>(scratch6);
    if (neg) dst[0] = '-';
    for (uint32_t i = 0; i < nd; ++i) dst[neg + i] = tmp[nd - 1 - i];
    for (uint32_t i = len; i < 24; ++i) dst[i] = 0;
    return len;
}
#endif  // __CUDACC__

}  // namespace rf
