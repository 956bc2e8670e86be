"""The row-sharded step's kernels on ONE GPU with `world` virtual owners whose buffers are all local: routing
(string keys and pre-hashed ids, staged and round-1 variants), owner-side pooling over gapped bags (ordered partials
and the red.global.add accumulate mode) and the combine -- against the numpy restatement of the sharded algorithm
(bit-exact for the ordered path) and the sequential oracle (re-association bound for the accumulate path).
The N-GPU transport itself is covered by tests/test_sharded_gpu.py (needs >= 2 GPUs)."""
import numpy as np
import pytest
import torch

import oracle
from recommendflow_b200.bag_ops import hash_strings
from recommendflow_b200.sharded import CudaShardOps, shard_rows
from recommendflow_b200.strings import StringColumn
from tests.shard_util import rank_batch, sharded_reference

pytestmark = pytest.mark.gpu


def _route_all(ops, keys, bag_offsets, N, B, W, max_keys, ids_ws):
    rows = torch.zeros(W, max_keys, dtype=torch.int64, device="cuda")
    beg = torch.zeros(W, B, dtype=torch.int32, device="cuda")
    end = torch.zeros(W, B, dtype=torch.int32, device="cuda")
    ops.route_tiles(keys, N, "", None, ids_ws, bag_offsets, 0, B, W, [rows[g].data_ptr() for g in range(W)],
                    [beg[g].data_ptr() for g in range(W)], [end[g].data_ptr() for g in range(W)])
    return rows, beg, end


@pytest.mark.parametrize("world", [2, 8, 5])
@pytest.mark.parametrize("prehashed", [False, True])
def test_virtual_owner_sharded_step(world, prehashed):
    W, N, D, B, max_len = world, 100003, 64, 700, 230
    rng = np.random.default_rng(world)
    full = rng.uniform(-0.05, 0.05, size=(N, D)).astype(np.float32)
    arena, offs, bag = rank_batch(3, B, max_len)
    bag[5] = bag[4]                                           # bag 4 is empty
    bag = np.maximum.accumulate(bag)
    col = StringColumn.from_arena(arena, offs, (B, None), bag).to("cuda")
    ids = oracle.hash_strings(arena, offs, N, "", None)
    ops = CudaShardOps()
    n = col.n_items
    ids_ws = torch.full((n,), -1, dtype=torch.int64, device="cuda")
    keys = torch.from_numpy(ids).cuda() if prehashed else col
    rows, beg, end = _route_all(ops, keys, col.bag_offsets, N, B, W, n, None if prehashed else ids_ws)
    torch.cuda.synchronize()
    if not prehashed:
        assert np.array_equal(ids_ws.cpu().numpy(), ids)      # the fused hash equals the oracle's
    # every owner's run of every bag: the owner-local rows of that owner's keys, in key order
    rows_h, beg_h, end_h = rows.cpu().numpy(), beg.cpu().numpy(), end.cpu().numpy()
    for b in (0, 4, 5, 17, B - 1):
        k = ids[bag[b]:bag[b + 1]]
        for g in range(W):
            assert np.array_equal(rows_h[g, beg_h[g, b]:end_h[g, b]], k[k % W == g] // W), (b, g)
    # owner side: every virtual owner pools its shard for this one source; then combine in rank order
    shards = [torch.from_numpy(np.ascontiguousarray(full[g::W])).cuda() for g in range(W)]
    for combiner in ("avg", "sum", "max"):
        partials = torch.empty(W, B, D, dtype=torch.float32, device="cuda")
        for g in range(W):
            assert shards[g].shape[0] == shard_rows(N, g, W)
            ops.pool(shards[g], [rows[g]], [beg[g]], [partials[g]], B, "sum" if combiner == "avg" else combiner,
                     max(1, n // W), [end[g]])
        out = torch.empty(B, D, dtype=torch.float32, device="cuda")
        ops.combine(partials, W, B, D, combiner, 0, col.bag_offsets, out)
        want = sharded_reference(ids, bag, full, W, combiner)
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32)), combiner
        if combiner == "max":
            continue
        # accumulate mode: all owners add into ONE zeroed buffer, then the finishing pass (combine with world = 1)
        acc = torch.zeros(1, B, D, dtype=torch.float32, device="cuda")
        for g in range(W):
            ops.pool(shards[g], [rows[g]], [beg[g]], [acc[0]], B, "sum", max(1, n // W), [end[g]], accumulate=True,
                     max_ctas_per_sm=2 if g % 2 else 0)
        ops.combine(acc, 1, B, D, combiner, 0, col.bag_offsets, out)
        seq = oracle.bag_pool(ids, full, combiner, bag_offsets=bag)
        assert float(np.abs(out.cpu().numpy() - seq).max()) <= max_len * 0.05 * 2.0 ** -21
        assert not out[4].any()                               # the empty bag stays 0


def test_accumulate_flag_rules():
    from recommendflow_b200 import _native as nat
    from recommendflow_b200.bag_ops import FieldCall, bag_forward
    w = torch.zeros(10, 4, device="cuda")
    ids = torch.zeros(1, 8, dtype=torch.int64, device="cuda")
    out = torch.zeros(4, 4, device="cuda")
    ok = FieldCall([(w, 10, None)], 4, "sum", ids=ids, out=out, bag_len=2, flags=nat.FIELD_ACCUMULATE)
    plain = FieldCall([(w, 10, None)], 4, "sum", ids=ids, out=out, bag_len=2)
    with pytest.raises(ValueError):
        bag_forward([ok, plain], 4)                          # the flag is launch-wide
    with pytest.raises(ValueError):
        bag_forward([FieldCall([(w, 10, None)], 4, "max", ids=ids, out=out, bag_len=2, flags=nat.FIELD_ACCUMULATE)], 4)
    w.fill_(1.0)
    bag_forward([ok], 4)
    bag_forward([ok], 4)
    assert torch.equal(out, torch.full_like(out, 4.0))        # two launches of 2 keys x 1.0 accumulated
