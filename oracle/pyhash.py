"""Pure-Python restatement of the two TF string hashes (ORACLE -- test infrastructure only).

Independent of oracle/rf_oracle.c on purpose (different language, big-int arithmetic with an
explicit 64-bit mask, byte slicing via int.from_bytes) so that the two agreeing on inputs
longer than 16 bytes -- for which no public known-answer vector exists --
is a meaningful cross-check.  Small cases only: this is a loop over Python ints.

Follows the published algorithms that the reference reaches through Keras `Hashing`
(/root/reference/backend/layers/preprocess_layers.py:89-90):
  salt=None     -> tf.strings.to_hash_bucket_fast   -> farmhashna::Hash64 ("Fingerprint64")
  salt=[k0,k1]  -> tf.strings.to_hash_bucket_strong -> highwayhash SipHash-2-4
"""
M64 = (1 << 64) - 1
_K0 = 0xC3A5C85C97CB3127
_K1 = 0xB492B66FBE98F273
_K2 = 0x9AE16A3B2F90404F


def _le(b, i, w):
    return int.from_bytes(b[i:i + w], "little")


def _rotr(v, s):
    return v if s == 0 else ((v >> s) | (v << (64 - s))) & M64


def _mix(v):
    return v ^ (v >> 47)


def _h16(u, v, mul):
    a = ((u ^ v) * mul) & M64
    a ^= a >> 47
    b = ((v ^ a) * mul) & M64
    b ^= b >> 47
    return (b * mul) & M64


def _weak(b, p, a, bb):
    w, x, y, z = _le(b, p, 8), _le(b, p + 8, 8), _le(b, p + 16, 8), _le(b, p + 24, 8)
    a = (a + w) & M64
    bb = _rotr((bb + a + z) & M64, 21)
    c = a
    a = (a + x + y) & M64
    bb = (bb + _rotr(a, 44)) & M64
    return (a + z) & M64, (bb + c) & M64


def fingerprint64(data: bytes) -> int:
    n = len(data)
    if n <= 16:
        if n >= 8:
            mul = (_K2 + 2 * n) & M64
            a = (_le(data, 0, 8) + _K2) & M64
            b = _le(data, n - 8, 8)
            c = (_rotr(b, 37) * mul + a) & M64
            d = ((_rotr(a, 25) + b) * mul) & M64
            return _h16(c, d, mul)
        if n >= 4:
            mul = (_K2 + 2 * n) & M64
            a = _le(data, 0, 4)
            return _h16((n + (a << 3)) & M64, _le(data, n - 4, 4), mul)
        if n > 0:
            y = (data[0] + (data[n >> 1] << 8)) & 0xFFFFFFFF
            z = (n + (data[n - 1] << 2)) & 0xFFFFFFFF
            return (_mix(((y * _K2) & M64) ^ ((z * _K0) & M64)) * _K2) & M64
        return _K2
    if n <= 32:
        mul = (_K2 + 2 * n) & M64
        a = (_le(data, 0, 8) * _K1) & M64
        b = _le(data, 8, 8)
        c = (_le(data, n - 8, 8) * mul) & M64
        d = (_le(data, n - 16, 8) * _K2) & M64
        return _h16((_rotr((a + b) & M64, 43) + _rotr(c, 30) + d) & M64,
                    (a + _rotr((b + _K2) & M64, 18) + c) & M64, mul)
    if n <= 64:
        mul = (_K2 + 2 * n) & M64
        a = (_le(data, 0, 8) * _K2) & M64
        b = _le(data, 8, 8)
        c = (_le(data, n - 8, 8) * mul) & M64
        d = (_le(data, n - 16, 8) * _K2) & M64
        y = (_rotr((a + b) & M64, 43) + _rotr(c, 30) + d) & M64
        z = _h16(y, (a + _rotr((b + _K2) & M64, 18) + c) & M64, mul)
        e = (_le(data, 16, 8) * mul) & M64
        f = _le(data, 24, 8)
        g = ((y + _le(data, n - 32, 8)) * mul) & M64
        h = ((z + _le(data, n - 24, 8)) * mul) & M64
        return _h16((_rotr((e + f) & M64, 43) + _rotr(g, 30) + h) & M64,
                    (e + _rotr((f + a) & M64, 18) + g) & M64, mul)
    seed = 81
    x = seed
    y = (seed * _K1 + 113) & M64
    z = (_mix((y * _K2 + 113) & M64) * _K2) & M64
    v = (0, 0)
    w = (0, 0)
    x = (x * _K2 + _le(data, 0, 8)) & M64
    pos = 0
    end = ((n - 1) // 64) * 64
    last64 = end + ((n - 1) & 63) - 63
    while True:
        x = (_rotr((x + y + v[0] + _le(data, pos + 8, 8)) & M64, 37) * _K1) & M64
        y = (_rotr((y + v[1] + _le(data, pos + 48, 8)) & M64, 42) * _K1) & M64
        x ^= w[1]
        y = (y + v[0] + _le(data, pos + 40, 8)) & M64
        z = (_rotr((z + w[0]) & M64, 33) * _K1) & M64
        v = _weak(data, pos, (v[1] * _K1) & M64, (x + w[0]) & M64)
        w = _weak(data, pos + 32, (z + w[1]) & M64, (y + _le(data, pos + 16, 8)) & M64)
        z, x = x, z
        pos += 64
        if pos == end:
            break
    mul = (_K1 + ((z & 0xFF) << 1)) & M64
    pos = last64
    w = ((w[0] + ((n - 1) & 63)) & M64, w[1])
    v = ((v[0] + w[0]) & M64, v[1])
    w = ((w[0] + v[0]) & M64, w[1])
    x = (_rotr((x + y + v[0] + _le(data, pos + 8, 8)) & M64, 37) * mul) & M64
    y = (_rotr((y + v[1] + _le(data, pos + 48, 8)) & M64, 42) * mul) & M64
    x ^= (w[1] * 9) & M64
    y = (y + v[0] * 9 + _le(data, pos + 40, 8)) & M64
    z = (_rotr((z + w[0]) & M64, 33) * mul) & M64
    v = _weak(data, pos, (v[1] * mul) & M64, (x + w[0]) & M64)
    w = _weak(data, pos + 32, (z + w[1]) & M64, (y + _le(data, pos + 16, 8)) & M64)
    z, x = x, z
    return _h16((_h16(v[0], w[0], mul) + ((_mix(y) * _K0) & M64) + z) & M64,
                (_h16(v[1], w[1], mul) + x) & M64, mul)


def _rotl(v, s):
    return ((v << s) | (v >> (64 - s))) & M64


def siphash24(k0: int, k1: int, data: bytes) -> int:
    v = [k0 ^ 0x736F6D6570736575, k1 ^ 0x646F72616E646F6D,
         k0 ^ 0x6C7967656E657261, k1 ^ 0x7465646279746573]

    def rnd():
        v[0] = (v[0] + v[1]) & M64
        v[1] = _rotl(v[1], 13) ^ v[0]
        v[0] = _rotl(v[0], 32)
        v[2] = (v[2] + v[3]) & M64
        v[3] = _rotl(v[3], 16) ^ v[2]
        v[0] = (v[0] + v[3]) & M64
        v[3] = _rotl(v[3], 21) ^ v[0]
        v[2] = (v[2] + v[1]) & M64
        v[1] = _rotl(v[1], 17) ^ v[2]
        v[2] = _rotl(v[2], 32)

    n = len(data)
    full = n - (n % 8)
    for i in range(0, full, 8):
        m = _le(data, i, 8)
        v[3] ^= m
        rnd(); rnd()
        v[0] ^= m
    m = ((n & 0xFF) << 56) | int.from_bytes(data[full:], "little")
    v[3] ^= m
    rnd(); rnd()
    v[0] ^= m
    v[2] ^= 0xFF
    rnd(); rnd(); rnd(); rnd()
    return v[0] ^ v[1] ^ v[2] ^ v[3]


def keras_hashing(values, num_bins, mask_value=None, salt=None):
    """Keras `Hashing(num_bins, mask_value, salt)` on a flat list of str/bytes/int."""
    if isinstance(salt, int):
        salt = [salt, salt]
    bins = num_bins
    masking = mask_value is not None and num_bins > 1
    if masking:
        bins -= 1
    out = []
    for x in values:
        is_mask = masking and x == mask_value
        s = str(x).encode() if isinstance(x, int) else (x.encode() if isinstance(x, str) else bytes(x))
        h = siphash24(salt[0], salt[1], s) if salt is not None else fingerprint64(s)
        i = h % bins
        if masking:
            i = 0 if is_mask else i + 1
        out.append(i)
    return out
