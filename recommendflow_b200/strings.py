"""StringColumn: a [B, L] batch of byte strings as the kernels consume it.

The reference feeds Keras layers dense `tf.string` tensors, zero-padded with "" to the longest
list in the batch (FixedLenSequenceFeature(allow_missing=True, default_value=""),
/root/reference/backend/core/dataloader.py:33).  Here the same batch is one contiguous uint8
arena plus int32 offsets[B*L + 1]; a pad is simply an empty string.  The arena carries 16
bytes of slack so the device can fetch whole words past the last key.
"""
import numpy as np
import torch

ARENA_SLACK = 16


class StringColumn(object):
    def __init__(self, data, offsets, shape, bag_offsets=None):
        """data: uint8 tensor (arena + slack); offsets: int32 [n+1]; shape: (B, L) dense, or
        (B, None) with bag_offsets int32 [B+1] for a jagged column."""
        self.data = data
        self.offsets = offsets
        self.shape = tuple(shape)
        self.bag_offsets = bag_offsets

    @property
    def n_items(self):
        return int(self.offsets.numel()) - 1

    @property
    def device(self):
        return self.data.device

    @property
    def nbytes(self):
        return int(self.data.numel()) - ARENA_SLACK

    def to(self, device, non_blocking=False):
        bo = None if self.bag_offsets is None else self.bag_offsets.to(device, non_blocking=non_blocking)
        return StringColumn(self.data.to(device, non_blocking=non_blocking),
                            self.offsets.to(device, non_blocking=non_blocking), self.shape, bo)

    def pin_memory(self):
        bo = None if self.bag_offsets is None else self.bag_offsets.pin_memory()
        return StringColumn(self.data.pin_memory(), self.offsets.pin_memory(), self.shape, bo)

    # ---- constructors ----------------------------------------------------------------------
    @staticmethod
    def from_arena(arena, offsets, shape, bag_offsets=None):
        """numpy uint8 arena + int offsets -> host StringColumn (adds the slack)."""
        arena = np.ascontiguousarray(arena, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        if offsets.size and int(offsets[-1]) != arena.size:
            raise ValueError("offsets[-1] must equal the arena size")
        buf = np.zeros(arena.size + ARENA_SLACK, dtype=np.uint8)
        buf[:arena.size] = arena
        bo = None if bag_offsets is None else torch.from_numpy(np.ascontiguousarray(bag_offsets, dtype=np.int32))
        return StringColumn(torch.from_numpy(buf), torch.from_numpy(offsets), shape, bo)

    @staticmethod
    def from_lists(rows, jagged=False):
        """rows: list (batch) of list of str/bytes (or a single str per row).

        Dense (default): every row is padded with "" to the longest row, exactly what
        tf.io.parse_example produces for the reference.  jagged=True keeps the true lengths."""
        norm = []
        for r in rows:
            if isinstance(r, (str, bytes)):
                r = [r]
            norm.append([x.encode() if isinstance(x, str) else bytes(x) for x in r])
        B = len(norm)
        if jagged:
            flat = [x for r in norm for x in r]
            bag = np.zeros(B + 1, dtype=np.int32)
            bag[1:] = np.cumsum([len(r) for r in norm])
            shape = (B, None)
        else:
            L = max((len(r) for r in norm), default=0)
            flat = [x for r in norm for x in (r + [b""] * (L - len(r)))]
            bag = None
            shape = (B, L)
        offs = np.zeros(len(flat) + 1, dtype=np.int64)
        if flat:
            offs[1:] = np.cumsum([len(x) for x in flat])
        if offs[-1] >= 2**31:
            raise ValueError("string arena of one column must stay below 2 GiB")
        arena = np.frombuffer(b"".join(flat), dtype=np.uint8)
        return StringColumn.from_arena(arena, offs.astype(np.int32), shape, bag)

    @staticmethod
    def from_numpy(arr):
        """2-D numpy array of str/bytes objects ([B, L], already padded)."""
        arr = np.asarray(arr)
        if arr.ndim == 1:
            arr = arr[:, None]
        return StringColumn.from_lists([list(row) for row in arr.tolist()])

    def tolist(self):
        data = self.data.cpu().numpy().tobytes()
        offs = self.offsets.cpu().numpy()
        return [data[offs[i]:offs[i + 1]] for i in range(self.n_items)]
