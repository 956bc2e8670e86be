"""recommendflow_b200.graphs.GraphedCall: a forward of the layer API recorded into one CUDA graph over static buffers."""
import numpy as np
import pytest
import torch

from recommendflow_b200 import _native as nat
from recommendflow_b200 import dense_ops
from recommendflow_b200.backend.layers.preprocess_layers import DoubleHashingEmbedding
from recommendflow_b200.graphs import GraphedCall
from recommendflow_b200.strings import StringColumn

pytestmark = pytest.mark.gpu


def test_graphed_call_replays_the_forward_on_refilled_buffers():
    """Fused hash + gather + pool (descriptors uploaded by a kernel node from their pinned slot) followed by the tcgen05 Dense
    kernel, recorded once; replays on refilled key buffers equal the eager calls bit for bit and issue no host-side launch."""
    rng = np.random.default_rng(5)
    B, L, N, D = 768, 3, 5000, 8
    layer = DoubleHashingEmbedding(num_bins=N, output_dim=D, seeds=[11, 12], combiner="sum", mask_value="", mask_zero=True, name="hashing_x")
    layer.set_weights([rng.uniform(-0.05, 0.05, size=(N, D)).astype(np.float32) for _ in range(2)])
    wt = torch.from_numpy((rng.standard_normal((32, 2 * D)) * 0.3).astype(np.float32)).cuda()
    bias = torch.from_numpy(rng.uniform(-0.1, 0.1, 32).astype(np.float32)).cuda()

    def batch():                                                     # fixed-width keys: every batch fills the same buffers
        return StringColumn.from_lists([[f"k{rng.integers(0, 10**6):06d}" for _ in range(L)] for _ in range(B)]).to("cuda")

    batches = [batch() for _ in range(3)]
    static = StringColumn(batches[0].data.clone(), batches[0].offsets.clone(), batches[0].shape)

    def forward(col):
        return dense_ops.dense_forward(layer(col), wt, bias, "selu")

    step = GraphedCall(lambda: forward(static))
    for b in batches + batches[:1]:
        static.data.copy_(b.data)
        static.offsets.copy_(b.offsets)
        before = nat.launch_count()
        got = step().clone()
        assert nat.launch_count() == before, "a replay goes through no host-side launch"
        assert torch.equal(got, forward(b))
    step.release()
