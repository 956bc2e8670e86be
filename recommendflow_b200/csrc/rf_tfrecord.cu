// rf_tfrecord.cu -- host-side input codec behind include/rf_tfrecord.h (no device code; it is a .cu only so
// that the one nvcc invocation of build.py picks it up).  See the header for the wire formats and for what it
// replaces in the reference (tf.io.parse_example over utils/make_tfrecord.py's files).
#include <string.h>

#include <string>
#include <vector>

#include "../../include/rf_tfrecord.h"
#include "rf_common.h"

namespace rf {
namespace {

// ---- CRC32C (Castagnoli, reflected 0x82F63B78), slicing-by-8 --------------------------------------
struct CrcTables {
    uint32_t t[8][256];
    CrcTables() {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
            t[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; ++i)
            for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xff];
    }
};

const CrcTables &crc_tables() {
    static const CrcTables tables;
    return tables;
}

uint32_t crc32c(const uint8_t *p, int64_t n) {
    const CrcTables &T = crc_tables();
    uint32_t crc = 0xFFFFFFFFu;
    while (n >= 8) {
        uint32_t lo, hi;
        memcpy(&lo, p, 4);
        memcpy(&hi, p + 4, 4);
        lo ^= crc;
        crc = T.t[7][lo & 0xff] ^ T.t[6][(lo >> 8) & 0xff] ^ T.t[5][(lo >> 16) & 0xff] ^ T.t[4][lo >> 24] ^
              T.t[3][hi & 0xff] ^ T.t[2][(hi >> 8) & 0xff] ^ T.t[1][(hi >> 16) & 0xff] ^ T.t[0][hi >> 24];
        p += 8;
        n -= 8;
    }
    while (n-- > 0) crc = T.t[0][(crc ^ *p++) & 0xff] ^ (crc >> 8);
    return crc ^ 0xFFFFFFFFu;
}

uint32_t mask_crc(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + 0xA282EAD8u; }

// ---- protobuf wire primitives ------------------------------------------------------------------------
struct Span {
    const uint8_t *p, *end;
    bool empty() const { return p >= end; }
};

bool read_varint(Span &s, uint64_t &v) {
    v = 0;
    for (int shift = 0; shift < 64 && s.p < s.end; shift += 7) {
        const uint8_t b = *s.p++;
        v |= (uint64_t)(b & 0x7f) << shift;
        if (!(b & 0x80)) return true;
    }
    return false;
}

// advance over one field's payload given its wire type; length-delimited payloads are returned in `sub`
bool read_field(Span &s, uint32_t &field, uint32_t &wire, Span &sub, uint64_t &scalar) {
    uint64_t tag;
    if (!read_varint(s, tag)) return false;
    field = (uint32_t)(tag >> 3);
    wire = (uint32_t)(tag & 7);
    sub.p = sub.end = nullptr;
    scalar = 0;
    switch (wire) {
        case 0: return read_varint(s, scalar);
        case 1:
            if (s.end - s.p < 8) return false;
            memcpy(&scalar, s.p, 8);
            s.p += 8;
            return true;
        case 2: {
            uint64_t n;
            if (!read_varint(s, n) || n > (uint64_t)(s.end - s.p)) return false;
            sub.p = s.p;
            sub.end = s.p + n;
            s.p += n;
            return true;
        }
        case 5: {
            if (s.end - s.p < 4) return false;
            uint32_t v32;
            memcpy(&v32, s.p, 4);
            scalar = v32;
            s.p += 4;
            return true;
        }
        default: return false;
    }
}

// ---- key -> column map (open addressing, FNV-1a) -----------------------------------------------------------
struct KeyMap {
    std::vector<int> slots;
    uint32_t mask;
    const rf_example_column *cols;

    static uint32_t hash(const uint8_t *p, size_t n) {
        uint32_t h = 2166136261u;
        for (size_t i = 0; i < n; ++i) h = (h ^ p[i]) * 16777619u;
        return h;
    }
    void build(const rf_example_column *c, int n) {
        cols = c;
        uint32_t cap = 8;
        while (cap < (uint32_t)n * 2u) cap <<= 1;
        slots.assign(cap, -1);
        mask = cap - 1;
        for (int i = 0; i < n; ++i) {
            uint32_t pos = hash(reinterpret_cast<const uint8_t *>(c[i].name), (size_t)c[i].name_len) & mask;
            while (slots[pos] >= 0) pos = (pos + 1) & mask;
            slots[pos] = i;
        }
    }
    int find(const uint8_t *p, size_t n) const {
        uint32_t pos = hash(p, n) & mask;
        while (slots[pos] >= 0) {
            const rf_example_column &c = cols[slots[pos]];
            if ((size_t)c.name_len == n && memcmp(c.name, p, n) == 0) return slots[pos];
            pos = (pos + 1) & mask;
        }
        return -1;
    }
};

struct Cursor {     // running write positions of one column during the fill pass
    int64_t values, bytes;
};

const char *kind_name(int kind) { return kind == RF_TFR_BYTES ? "bytes_list" : kind == RF_TFR_FLOAT ? "float_list" : "int64_list"; }

// decode one Feature payload into column c (count or fill); returns the number of values or -1
int64_t decode_feature(Span feat, rf_example_column &c, Cursor &cur, bool fill, int64_t rec) {
    int64_t n_values = 0;
    while (!feat.empty()) {
        uint32_t field, wire;
        Span list;
        uint64_t scalar;
        if (!read_field(feat, field, wire, list, scalar)) return -1;
        if (wire != 2 || field < 1 || field > 3) continue;           // unknown member: skip
        const int kind = (int)field - 1;                             // 1 bytes_list, 2 float_list, 3 int64_list
        if (kind != c.kind) {
            set_error(RF_ERR_INVALID, "record %lld, feature %.*s: expected %s, found %s", (long long)rec, c.name_len, c.name,
                      kind_name(c.kind), kind_name(kind));
            return -2;
        }
        while (!list.empty()) {
            uint32_t f2, w2;
            Span payload;
            uint64_t sc;
            if (!read_field(list, f2, w2, payload, sc)) return -1;
            if (f2 != 1) continue;
            if (kind == RF_TFR_BYTES) {
                if (w2 != 2) return -1;
                const int64_t n = payload.end - payload.p;
                if (fill) {
                    c.value_offsets[cur.values] = (int32_t)cur.bytes;
                    memcpy(c.bytes_out + cur.bytes, payload.p, (size_t)n);
                }
                cur.bytes += n;
                cur.values += 1;
                n_values += 1;
            } else if (kind == RF_TFR_FLOAT) {
                if (w2 == 2) {                                       // packed
                    const int64_t n = (payload.end - payload.p) / 4;
                    if (fill) memcpy(c.floats_out + cur.values, payload.p, (size_t)n * 4);
                    cur.values += n;
                    n_values += n;
                } else if (w2 == 5) {
                    if (fill) {
                        const uint32_t bits = (uint32_t)sc;
                        memcpy(c.floats_out + cur.values, &bits, 4);
                    }
                    cur.values += 1;
                    n_values += 1;
                } else {
                    return -1;
                }
            } else {
                if (w2 == 2) {                                       // packed varints
                    while (!payload.empty()) {
                        uint64_t v;
                        if (!read_varint(payload, v)) return -1;
                        if (fill) c.ints_out[cur.values] = (int64_t)v;
                        cur.values += 1;
                        n_values += 1;
                    }
                } else if (w2 == 0) {
                    if (fill) c.ints_out[cur.values] = (int64_t)sc;
                    cur.values += 1;
                    n_values += 1;
                } else {
                    return -1;
                }
            }
        }
    }
    return n_values;
}

}  // namespace
}  // namespace rf

using namespace rf;

extern "C" {

uint32_t rf_crc32c(const uint8_t *data, int64_t n) { return crc32c(data, n); }

uint32_t rf_masked_crc32c(const uint8_t *data, int64_t n) { return mask_crc(crc32c(data, n)); }

int rf_tfrecord_index(const uint8_t *buf, int64_t len, int verify_crc, int64_t max_records, int64_t *rec_offsets, int64_t *rec_lens,
                      int64_t *n_records) {
    if (len < 0 || (len > 0 && !buf) || !n_records) return set_error(RF_ERR_INVALID, "rf_tfrecord_index: bad arguments");
    int64_t pos = 0, n = 0;
    while (pos < len) {
        if (len - pos < 12) return set_error(RF_ERR_INVALID, "truncated TFRecord header at byte %lld", (long long)pos);
        uint64_t length;
        uint32_t hcrc;
        memcpy(&length, buf + pos, 8);
        memcpy(&hcrc, buf + pos + 8, 4);
        if (length > (uint64_t)(len - pos - 12) || (uint64_t)(len - pos - 12) - length < 4)
            return set_error(RF_ERR_INVALID, "truncated TFRecord body at byte %lld (record %lld)", (long long)pos, (long long)n);
        const uint8_t *body = buf + pos + 12;
        if (verify_crc) {
            uint32_t bcrc;
            memcpy(&bcrc, body + length, 4);
            if (mask_crc(crc32c(buf + pos, 8)) != hcrc || mask_crc(crc32c(body, (int64_t)length)) != bcrc)
                return set_error(RF_ERR_INVALID, "TFRecord CRC mismatch in record %lld", (long long)n);
        }
        if (rec_offsets || rec_lens) {
            if (n >= max_records) return set_error(RF_ERR_INVALID, "more than max_records = %lld records", (long long)max_records);
            if (rec_offsets) rec_offsets[n] = pos + 12;
            if (rec_lens) rec_lens[n] = (int64_t)length;
        }
        ++n;
        pos += 12 + (int64_t)length + 4;
    }
    *n_records = n;
    return RF_OK;
}

int rf_example_parse_columns(const uint8_t *buf, const int64_t *rec_offsets, const int64_t *rec_lens, int64_t n_records,
                             rf_example_column *cols, int n_cols, int fill) {
    if (n_records < 0 || n_cols < 0 || (n_cols > 0 && !cols) || (n_records > 0 && (!buf || !rec_offsets || !rec_lens)))
        return set_error(RF_ERR_INVALID, "rf_example_parse_columns: bad arguments");
    for (int i = 0; i < n_cols; ++i) {
        rf_example_column &c = cols[i];
        if (!c.name || c.name_len < 0 || c.kind < RF_TFR_BYTES || c.kind > RF_TFR_INT64)
            return set_error(RF_ERR_INVALID, "column %d: bad name / kind", i);
        if (fill) {
            const bool ok = c.kind == RF_TFR_BYTES ? (c.value_offsets && (c.bytes_out || c.n_bytes == 0))
                          : c.kind == RF_TFR_FLOAT ? (c.floats_out || c.n_values == 0) : (c.ints_out || c.n_values == 0);
            if (!ok) return set_error(RF_ERR_INVALID, "column %.*s: output buffers missing for the fill pass", c.name_len, c.name);
        }
    }
    KeyMap map;
    map.build(cols, n_cols);
    std::vector<Cursor> cur((size_t)n_cols, Cursor{0, 0});
    std::vector<Span> found((size_t)n_cols);
    std::vector<int> touched;
    for (int64_t r = 0; r < n_records; ++r) {
        Span ex{buf + rec_offsets[r], buf + rec_offsets[r] + rec_lens[r]};
        touched.clear();
        while (!ex.empty()) {                                     // Example: field 1 = Features
            uint32_t field, wire;
            Span features;
            uint64_t sc;
            if (!read_field(ex, field, wire, features, sc)) return set_error(RF_ERR_INVALID, "record %lld: malformed Example", (long long)r);
            if (field != 1 || wire != 2) continue;
            while (!features.empty()) {                           // Features: field 1 = map entry
                Span entry;
                if (!read_field(features, field, wire, entry, sc)) return set_error(RF_ERR_INVALID, "record %lld: malformed Features", (long long)r);
                if (field != 1 || wire != 2) continue;
                Span key{nullptr, nullptr}, feat{nullptr, nullptr};
                bool has_feat = false;
                while (!entry.empty()) {                          // entry: 1 = key, 2 = Feature
                    Span sub;
                    if (!read_field(entry, field, wire, sub, sc)) return set_error(RF_ERR_INVALID, "record %lld: malformed map entry", (long long)r);
                    if (wire != 2) continue;
                    if (field == 1) key = sub;
                    if (field == 2) {
                        feat = sub;
                        has_feat = true;
                    }
                }
                if (!key.p) continue;
                const int ci = map.find(key.p, (size_t)(key.end - key.p));
                if (ci < 0) continue;
                if (!found[(size_t)ci].p && !(found[(size_t)ci].end)) touched.push_back(ci);
                // an entry without a value is an empty Feature; keep a non-null marker either way
                static const uint8_t kNone = 0;
                found[(size_t)ci] = has_feat && feat.p ? feat : Span{&kNone, &kNone};
            }
        }
        for (int i = 0; i < n_cols; ++i)
            if (cols[i].row_counts) cols[i].row_counts[r] = 0;
        for (int ci : touched) {
            rf_example_column &c = cols[ci];
            const int64_t n = decode_feature(found[(size_t)ci], c, cur[(size_t)ci], fill != 0, r);
            if (n == -2) return RF_ERR_INVALID;
            if (n < 0) return set_error(RF_ERR_INVALID, "record %lld, feature %.*s: malformed Feature", (long long)r, c.name_len, c.name);
            if (c.row_counts) c.row_counts[r] = (int32_t)n;
            found[(size_t)ci] = Span{nullptr, nullptr};
        }
    }
    for (int i = 0; i < n_cols; ++i) {
        rf_example_column &c = cols[i];
        if (fill) {
            if (cur[(size_t)i].values != c.n_values || (c.kind == RF_TFR_BYTES && cur[(size_t)i].bytes != c.n_bytes))
                return set_error(RF_ERR_INVALID, "column %.*s: the fill pass found other sizes than the count pass", c.name_len, c.name);
            if (c.kind == RF_TFR_BYTES) c.value_offsets[c.n_values] = (int32_t)c.n_bytes;
        } else {
            c.n_values = cur[(size_t)i].values;
            c.n_bytes = cur[(size_t)i].bytes;
            if (c.kind == RF_TFR_BYTES && c.n_bytes > INT32_MAX)
                return set_error(RF_ERR_INVALID, "column %.*s: %lld bytes exceed the int32 offsets of one arena", c.name_len, c.name,
                                 (long long)c.n_bytes);
        }
    }
    return RF_OK;
}

}  // extern "C"
