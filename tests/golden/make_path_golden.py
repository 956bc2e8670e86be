"""Generate tests/golden/path_golden.json: seeded inputs and their outputs along the hot path, produced by the
CPU oracle AFTER it passed its known-answer tests (tests/test_oracle_kat.py).

    python tests/golden/make_path_golden.py

The reference cannot run here (TensorFlow is absent), so these vectors are not reference outputs;
they freeze the KAT-pinned oracle so that (a) a later change to the oracle or to oracle/pyhash.py cannot drift
silently and (b) the CUDA path is compared with committed numbers, not only with a checker built in the same
session.  Content:
  * strings of every FarmHash length branch (0, 1-3, 4-7, 8-16, 17-32, 33-64, 65+) with Fingerprint64,
    SipHash-2-4 under the reference's seeds, and the Keras `Hashing` bucket for three configurations;
  * int64 keys (hashed as decimal strings), same outputs;
  * "C1 substitute": the `app_id` feature of conf/base_conf.yaml (N = 3000, D = 16, sum, seeds [2022, 2023]) as a
    [6, 3] padded batch through DoubleHashingEmbedding with tables of a recorded seed: ids and pooled output bits.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402


def main():
    rng = np.random.default_rng(20261018)
    lengths = [0, 1, 2, 3, 4, 5, 7, 8, 9, 12, 15, 16, 17, 20, 24, 31, 32, 33, 40, 48, 63, 64, 65, 66, 100, 127, 128, 129, 200, 257, 1000]
    strings = [bytes(rng.integers(0, 256, size=n, dtype=np.uint8)) for n in lengths]
    strings += [b"app_id_0123456789", b"f00_1234567", b"Hello", b"World", "用户".encode(), b"-1", b"0"]
    arena, offs = oracle.encode_strings(strings)
    out = {"strings_hex": [s.hex() for s in strings],
           "fingerprint64": [str(oracle.fingerprint64(s)) for s in strings],
           "siphash24_2022": [str(oracle.siphash24(2022, 2022, s)) for s in strings],
           "siphash24_2023": [str(oracle.siphash24(2023, 2023, s)) for s in strings],
           "hashing": []}
    configs = [(3000, "", [2022, 2022]), (1000000, "", None), (2**32 - 1, None, [2023, 2023]), (100000, "", 2023)]
    for num_bins, mask, salt in configs:
        out["hashing"].append({"num_bins": num_bins, "mask_value": mask, "salt": salt,
                               "ids": oracle.hash_strings(arena, offs, num_bins, mask, salt).tolist()})
    ints = [0, 1, -1, 7, 10, 99, -100, 2022, 123456789, -987654321, 2**31 - 1, -2**31, 2**63 - 1, -2**63, 10**18]
    out["ints"] = [str(v) for v in ints]
    out["int_hashing"] = []
    for num_bins, mask, salt in [(3000, None, [2022, 2022]), (1000000, 0, None), (97, None, None)]:
        out["int_hashing"].append({"num_bins": num_bins, "mask_value": mask, "salt": salt,
                                   "ids": oracle.hash_ints(np.array(ints, dtype=np.int64), num_bins, mask, salt).tolist()})
    # C1 substitute: base_conf.yaml's app_id feature
    N, D, B, L = 3000, 16, 6, 3
    rows = [["com.tencent.mm", "com.ss.android.ugc.aweme", ""], ["com.eg.android.AlipayGphone", "", ""], ["", "", ""],
            ["a", "b", "c"], ["com.tencent.mm", "com.tencent.mm", "com.tencent.mm"], ["x" * 70, "y" * 33, "z" * 17]]
    flat = [x for r in rows for x in r]
    a2, o2 = oracle.encode_strings(flat)
    tables = [np.random.default_rng(7 + t).uniform(-0.05, 0.05, size=(N, D)).astype(np.float32) for t in range(2)]
    pooled = oracle.hashed_bag_forward(a2, o2, B, L, tables, [N, N], [2022, 2023], "sum")
    avg = oracle.hashed_bag_forward(a2, o2, B, L, tables, [N, N], [2022, 2023], "avg")
    out["app_id"] = {"num_bins": N, "dim": D, "seeds": [2022, 2023], "rows": rows, "table_rng_seeds": [7, 8],
                     "table_init": "numpy default_rng(seed).uniform(-0.05, 0.05, (N, D)).astype(float32)",
                     "ids1": oracle.hash_strings(a2, o2, N, "", 2022).tolist(), "ids2": oracle.hash_strings(a2, o2, N, "", 2023).tolist(),
                     "pooled_sum_bits": pooled.view(np.uint32).tolist(), "pooled_avg_bits": avg.view(np.uint32).tolist()}
    with open(os.path.join(HERE, "path_golden.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    print("wrote", os.path.join(HERE, "path_golden.json"), os.path.getsize(os.path.join(HERE, "path_golden.json")), "bytes")


if __name__ == "__main__":
    main()
