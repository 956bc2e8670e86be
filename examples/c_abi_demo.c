/* c_abi_demo.c -- the drop-in boundary used from plain C: no Python, no torch, only librf_b200.so + the CUDA runtime.
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include examples/c_abi_demo.c \
 *       -L recommendflow_b200 -lrf_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/recommendflow_b200 -o c_abi_demo
 *
 * It hashes the strings of the Keras `Hashing` docstring example (Hashing(num_bins=3, mask_value="") on
 * ["A", "B", "", "C", "D"] -> [1, 1, 0, 2, 2]) with rf_hash_strings, then pools them through a 3 x 8 table with
 * rf_bag_forward as one bag per string, and checks both against the expected values.
 * Exit codes: 0 = all checks passed, 1 = a check failed, 77 = no CUDA device (nothing was computed; there is no CPU
 * fallback).  tests/test_abi_cpu.py compiles and links it on every run; executing it needs a GPU box.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "rf_b200.h"

#define CHECK_CUDA(x)                                                          \
    do {                                                                       \
        cudaError_t e_ = (x);                                                  \
        if (e_ != cudaSuccess) {                                               \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));           \
            return 1;                                                          \
        }                                                                      \
    } while (0)

int main(void) {
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        fprintf(stderr, "no CUDA device: librf_b200 has no CPU fallback\n");
        return 77;
    }
    /* arena + offsets of ["A", "B", "", "C", "D"]; 16 bytes of slack after the last key */
    const unsigned char arena[4 + 16] = {'A', 'B', 'C', 'D'};
    const int32_t offsets[6] = {0, 1, 2, 2, 3, 4};
    const int64_t want_ids[5] = {1, 1, 0, 2, 2};
    unsigned char *d_bytes;
    int32_t *d_offsets;
    int64_t *d_ids;
    CHECK_CUDA(cudaMalloc((void **)&d_bytes, sizeof arena));
    CHECK_CUDA(cudaMalloc((void **)&d_offsets, sizeof offsets));
    CHECK_CUDA(cudaMalloc((void **)&d_ids, sizeof want_ids));
    CHECK_CUDA(cudaMemcpy(d_bytes, arena, sizeof arena, cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(d_offsets, offsets, sizeof offsets, cudaMemcpyHostToDevice));
    if (rf_hash_strings(d_bytes, d_offsets, 5, 3, RF_MASK_EMPTY_STRING, 0, 0, 0, d_ids, NULL) != RF_OK) {
        fprintf(stderr, "rf_hash_strings: %s\n", rf_last_error());
        return 1;
    }
    int64_t ids[5];
    CHECK_CUDA(cudaMemcpy(ids, d_ids, sizeof ids, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int i = 0; i < 5; ++i) {
        printf("id[%d] = %lld (expected %lld)\n", i, (long long)ids[i], (long long)want_ids[i]);
        bad += ids[i] != want_ids[i];
    }

    /* the same keys, one bag each, through a 3 x 8 table: out[i] must be table row ids[i] */
    float table[3][8] = {{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, {1.f, 2.f, 3.f, 4.f, 5.f, 6.f, 7.f, 8.f},
                         {-1.f, -2.f, -3.f, -4.f, -5.f, -6.f, -7.f, -8.f}};
    float *d_table, *d_out, out[5][8];
    CHECK_CUDA(cudaMalloc((void **)&d_table, sizeof table));
    CHECK_CUDA(cudaMalloc((void **)&d_out, sizeof out));
    CHECK_CUDA(cudaMemcpy(d_table, table, sizeof table, cudaMemcpyHostToDevice));
    rf_field_desc f;
    memset(&f, 0, sizeof f);
    f.bytes = d_bytes;
    f.str_offsets = d_offsets;
    f.bag_len = 1;
    f.n_tables = 1;
    f.tables[0].weights = d_table;
    f.tables[0].num_bins = 3;
    f.dim = 8;
    f.combiner = RF_COMBINER_SUM;
    f.mask_mode = RF_MASK_EMPTY_STRING;
    f.out = d_out;
    f.out_stride = 8;
    if (rf_bag_forward(&f, 1, 5, NULL) != RF_OK) {
        fprintf(stderr, "rf_bag_forward: %s\n", rf_last_error());
        return 1;
    }
    CHECK_CUDA(cudaMemcpy(out, d_out, sizeof out, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 5; ++i)
        for (int c = 0; c < 8; ++c) bad += out[i][c] != table[want_ids[i]][c];
    printf("%s\n", bad ? "MISMATCH" : "ok: hash + gather + pool through the C-ABI");
    cudaFree(d_bytes);
    cudaFree(d_offsets);
    cudaFree(d_ids);
    cudaFree(d_table);
    cudaFree(d_out);
    return bad ? 1 : 0;
}
