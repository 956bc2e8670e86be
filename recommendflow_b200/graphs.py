"""CUDA-graph wrappers for the path's steps.

An eager step of the layer API pays the host for every launch: a C3 forward issues 13 kernels plus layer glue from Python, a
training step ~350 launches (after the round-2 host-side trims: 0.64 ms eager vs 0.61 ms replayed for the forward, 5.5 vs
4.9 ms for the training step; the gap grows with smaller batches and in pipelined end-to-end loops, where a replay leaves the
host free to copy the next batch).  Everything on the path is capture-safe (no host
synchronisation, descriptors of captured launches live in dedicated pinned slots and are uploaded by a kernel node), so a
step can be recorded once over STATIC buffers and replayed with one launch:

    step = GraphedCall(lambda: model.loss_fun(y, *model.towers(keys, behaviour)))     # keys, y, behaviour: device buffers
    keys.data.copy_(next_batch.data, non_blocking=True); ...                           # refill the same buffers in place
    loss = step()                                                                      # replay; `loss` is overwritten

`training.GraphedTrainStep` is the same for a whole optimisation step (it also moves the optimizers' step counters to the device).
"""
import torch

from . import _native as nat


class GraphedCall(object):
    """Record `fn()` (a closure over static CUDA tensors) into a CUDA graph.  `fn` runs `warmup` times eagerly first (launch plans,
    workspaces and descriptor pools are created outside the capture), then once under capture; `__call__` replays and returns
    the captured outputs (the same tensors every time)."""

    def __init__(self, fn, warmup=2, device=None):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = fn()

    def __call__(self):
        self.graph.replay()
        return self.outputs

    def release(self):
        """Drop the graph and hand its descriptor slots back (rf_release_captured_launches frees the slots of EVERY captured
        launch of the process: call it when no other recorded step is alive)."""
        self.graph = None
        torch.cuda.synchronize()
        nat.check(nat.lib().rf_release_captured_launches())
