"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/rf_b200.h
declares, and its host-side helpers are right.  No compute entry point is called here."""
import ctypes
import os
import random
import re

import pytest

from recommendflow_b200 import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for header in sorted(os.listdir(os.path.join(ROOT, "include"))):
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(rf_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = nat.lib()
    names = declared_symbols()
    assert "rf_bag_forward" in names and "rf_hash_strings" in names and len(names) >= 7
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rf_b200.h but not exported"
    assert lib.rf_abi_version() == 3


def test_struct_layout_matches_header_sizes():
    # rf_table_desc: ptr + i64 + 2*i32 + 2*u64 = 40 bytes; rf_field_desc packs without surprises
    assert ctypes.sizeof(nat.TableDesc) == 40
    assert ctypes.sizeof(nat.FieldDesc) == 6 * 8 + 8 + 8 + 2 * 40 + 16 + 8 + 8 + 8 + 8 + 32 + 8
    assert nat.FieldDesc.tables.offset == 64 and nat.FieldDesc.out.offset == 168


def test_fastmod_equals_modulo():
    lib = nat.lib()
    rng = random.Random(7)
    divisors = [1, 2, 3, 7, 10, 255, 256, 999, 2999, 99999, 999999, 1000000, 2**31 - 1, 2**31, 2**32 - 2,
                2**32 - 1, 2**32, 2**40 + 17, 2**63 - 25, 2**63, 2**64 - 1] + [rng.getrandbits(rng.randint(2, 64)) | 1 for _ in range(200)]
    edge = [0, 1, 2, 2**32 - 1, 2**32, 2**63 - 1, 2**63, 2**64 - 2, 2**64 - 1]
    for d in divisors:
        for x in edge + [rng.getrandbits(64) for _ in range(200)] + [d - 1, d, d + 1 if d < 2**64 - 1 else d, (2**64 - 1) // d * d]:
            x &= 2**64 - 1
            assert lib.rf_debug_fastmod(x, d) == x % d, (x, d)


def test_bad_arguments_fail_before_touching_the_gpu():
    lib = nat.lib()
    f = (nat.FieldDesc * 1)()
    rc = lib.rf_bag_forward(f, 1, 4, None)          # no key source set
    assert rc == nat.RF_ERR_INVALID and b"exactly one of" in lib.rf_last_error()
    with pytest.raises(ValueError):
        nat.check(rc)
    assert lib.rf_bag_forward(f, 0, 4, None) == nat.RF_OK    # nothing to do
    assert lib.rf_hash_strings(None, None, 0, 10, 0, 0, 0, 0, None, None) == nat.RF_OK


def test_salt_rules():
    assert nat.salt_to_key(None) == (0, 0, 0)
    assert nat.salt_to_key(133) == (1, 133, 133)
    assert nat.salt_to_key([2022, 2023]) == (1, 2022, 2023)
    with pytest.raises(ValueError):
        nat.salt_to_key([1, 2, 3])


def test_header_is_plain_c_and_ctypes_structs_match_it(tmp_path):
    # include/rf_b200.h must compile as C99 on its own (it is what a cgo / JNI / TF-op binding includes), and the
    # ctypes mirrors in _native.py must have exactly the C compiler's sizes and member offsets
    import subprocess
    src = tmp_path / "layout.c"
    checks = {"rf_table_desc": (nat.TableDesc, ["weights", "num_bins", "use_strong", "key0", "key1"]),
              "rf_field_desc": (nat.FieldDesc, ["bytes", "str_offsets", "int_values", "ids", "bag_offsets", "bag_ends", "n_items",
                                                "bag_len", "n_tables", "tables", "dim", "combiner", "mask_mode", "flags",
                                                "int_mask_value", "out", "out_stride", "ids_out", "mask_bytes", "mask_len"]),
              "rf_shard_ctx": (nat.ShardCtx, ["rank", "world", "max_batch", "max_keys", "dim", "peer_exchange", "peer_signals"]),
              "rf_vocab_desc": (nat.VocabDesc, ["term_bytes", "term_offsets", "term_ints", "slots", "capacity", "n_terms"]),
              "rf_adam_params": (nat.AdamParams, ["lr", "beta1", "beta2", "epsilon", "step", "lazy", "d_lr_t", "d_live_rows"]),
              "rf_example_column": (nat.ExampleColumn, ["name", "name_len", "kind", "n_values", "n_bytes", "row_counts", "bytes_out",
                                                        "value_offsets", "floats_out", "ints_out"]),
              "rf_adam_field": (nat.AdamField, ["ids", "bag_offsets", "n_keys", "bag_len", "combiner", "grad_out", "grad_stride",
                                                "table", "m", "v", "table_rows", "dim"])}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "rf_b200.h")}"', f'#include "{os.path.join(ROOT, "include", "rf_tfrecord.h")}"',
             "int main(void) {"]
    for cname, (_, members) in checks.items():
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for m in members:
            lines.append(f'  printf(" %zu", offsetof({cname}, {m}));')
        lines.append('  printf("\\n");')
    lines += ["  return 0;", "}"]
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).strip().splitlines()
    for line in out:
        cname, size, *offs = line.split()
        struct, members = checks[cname]
        assert ctypes.sizeof(struct) == int(size), cname
        assert [getattr(struct, m).offset for m in members] == [int(o) for o in offs], cname


def test_plain_c_program_links_against_the_library(tmp_path):
    # examples/c_abi_demo.c uses the boundary from C99 with nothing but the CUDA runtime: it must compile, link against
    # librf_b200.so and start.  Without a GPU it reports that and exits 77 (no CPU fallback); on a GPU box it checks a
    # Keras Hashing known answer and a gather through rf_bag_forward and exits 0.
    import subprocess
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    exe = str(tmp_path / "c_abi_demo")
    libdir = os.path.join(ROOT, "recommendflow_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
                           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-L", libdir, "-lrf_b200", "-L", os.path.join(cuda, "lib64"),
                           "-lcudart", f"-Wl,-rpath,{libdir}", "-o", exe])
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode in (0, 77), res.stdout + res.stderr
    if res.returncode == 77:
        assert "no CPU fallback" in res.stderr


def test_integration_doc_lists_every_exported_entry_point():
    # INTEGRATION.md is the binder's map from C-ABI symbols to the reference code they replace: no symbol may be missing from it
    import re
    declared = set()
    for h in ("rf_b200.h", "rf_tfrecord.h"):
        declared |= set(re.findall(r"\b(rf_[a-z0-9_]+)\s*\(", open(os.path.join(ROOT, "include", h)).read()))
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = sorted(n for n in declared if n not in doc)
    assert not missing, missing


def test_workspace_size_queries_are_host_only_and_consistent():
    # size queries never touch the device: usable on a CPU-only box (the driver's build check) and by a binder planning memory
    lib = nat.lib()
    # split-K workspace: only for few output tiles + a long contraction (dW = X^T dZ), never for the big forward products
    assert lib.rf_dense_tc_workspace_bytes(512, 8192, 256) > 0
    assert lib.rf_dense_tc_workspace_bytes(64, 409600, 192) > 0
    assert lib.rf_dense_tc_workspace_bytes(8192, 1888, 1024) == 0
    assert lib.rf_dense_tc_workspace_bytes(0, 64, 64) == 0
    # a split buffer holds whole 128-row tiles of the output, once per split
    w = lib.rf_dense_tc_workspace_bytes(512, 8192, 256)
    assert w % (512 * 256 * 4) == 0 and 2 <= w // (512 * 256 * 4) <= 64
    # tower training passes: 2 values per column per 256-row split
    assert lib.rf_tower_train_workspace_bytes(8192, 1888) == (8192 // 256) * 2 * 1888 * 4
    assert lib.rf_tower_train_workspace_bytes(1, 4) == 2 * 4 * 4
    # CE backward: slab + transpose + the two transposed operands + one partial; at most ~1 GiB of slab for any batch
    small, big = lib.rf_inbatch_ce_backward_tc_workspace_bytes(8192, 256), lib.rf_inbatch_ce_backward_tc_workspace_bytes(65536, 256)
    assert small >= 2 * 8192 * 8192 * 4 and big < 2 * (1 << 30) + 4 * 65536 * 256 * 4 + (1 << 20)


def test_integration_doc_ctypes_stub_matches_the_header():
    # the minimal ctypes binding printed in INTEGRATION.md is executable and lays its structs out like the C compiler does
    import ctypes as C
    import re
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"```python\n(import ctypes as C.*?)```", doc, re.S).group(1)
    code = code.replace('C.CDLL("recommendflow_b200/librf_b200.so")', f'C.CDLL("{os.path.join(ROOT, "recommendflow_b200", "librf_b200.so")}")')
    ns = {}
    exec(code, ns)
    assert C.sizeof(ns["FieldDesc"]) == C.sizeof(nat.FieldDesc) and C.sizeof(ns["TableDesc"]) == C.sizeof(nat.TableDesc)
    for name, _ in nat.FieldDesc._fields_:
        assert getattr(ns["FieldDesc"], name).offset == getattr(nat.FieldDesc, name).offset, name
