/* rf_tfrecord.h -- host-side C-ABI of the input codec: TFRecord framing + tf.train.Example columns.
 *
 * "Next" row 2 of SURVEY.md §8f.  The reference reads its training data with tf.data / tf.io.parse_example
 * (backend/core/dataloader.py:23-44, 541-578) from GZIP TFRecord files written by utils/make_tfrecord.py
 * (:87-119, 139-144).  These entry points turn the (already decompressed) byte stream of such a file into
 * exactly what the kernels consume -- per feature one byte arena + int32 offsets, or a flat float / int64
 * array, plus the number of values each record contributed -- without creating one object per value.
 * Plain host code (no CUDA calls); lives in librf_b200.so next to the kernels.
 *
 * Wire formats (public specifications):
 *   TFRecord   u64 length | u32 masked_crc32c(length) | bytes | u32 masked_crc32c(bytes)
 *              masked = rotr(crc32c, 15) + 0xa282ead8
 *   Example    Example{1: Features{1: map<string, Feature>}}
 *              Feature{oneof 1: BytesList{1: repeated bytes}, 2: FloatList{1: packed | repeated float},
 *                            3: Int64List{1: packed | repeated varint}}
 */
#ifndef RF_TFRECORD_H_
#define RF_TFRECORD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RF_TFR_BYTES 0
#define RF_TFR_FLOAT 1
#define RF_TFR_INT64 2

uint32_t rf_crc32c(const uint8_t *data, int64_t n);
uint32_t rf_masked_crc32c(const uint8_t *data, int64_t n);

/* Walk the framing of `buf[0, len)`: *n_records = number of records; when rec_offsets / rec_lens are   */
/* given (capacity max_records) they receive each record's payload position and length.  verify_crc    */
/* checks both CRCs of every record.  RF_ERR_INVALID (-1) on truncation / CRC mismatch / overflow of   */
/* max_records; message from rf_last_error().                                                          */
int rf_tfrecord_index(const uint8_t *buf, int64_t len, int verify_crc, int64_t max_records,
                      int64_t *rec_offsets, int64_t *rec_lens, int64_t *n_records);

/* One requested feature.  Pass 1 (fill = 0) sets n_values / n_bytes and row_counts; the caller then    */
/* allocates the outputs and runs pass 2 (fill = 1).  Values are emitted in record order, unpadded.     */
typedef struct rf_example_column {
    const char *name;        /* feature key (not NUL-terminated: name_len bytes)                        */
    int32_t name_len;
    int32_t kind;            /* RF_TFR_BYTES | RF_TFR_FLOAT | RF_TFR_INT64; another kind on the wire     */
                             /* is an error, as in tf.io.parse_example                                   */
    int64_t n_values;        /* out: values over all records                                            */
    int64_t n_bytes;         /* out (bytes kind): payload bytes over all records                        */
    int32_t *row_counts;     /* [n_records] out: values of each record (0 = feature absent)             */
    uint8_t *bytes_out;      /* bytes kind, pass 2: arena [n_bytes (+ slack the caller adds)]           */
    int32_t *value_offsets;  /* bytes kind, pass 2: [n_values + 1]                                      */
    float *floats_out;       /* float kind, pass 2: [n_values]                                          */
    int64_t *ints_out;       /* int64 kind, pass 2: [n_values]                                          */
} rf_example_column;

/* Decode the requested columns of n_records serialized Examples (payloads at buf + rec_offsets[i]).    */
/* A key that occurs twice in one Example keeps its last value (protobuf map semantics).                */
int rf_example_parse_columns(const uint8_t *buf, const int64_t *rec_offsets, const int64_t *rec_lens,
                             int64_t n_records, rf_example_column *cols, int n_cols, int fill);

#ifdef __cplusplus
}
#endif
#endif /* RF_TFRECORD_H_ */
