"""SMPL pose conditioning stage (reference ldm/modules/poses/poses.py:3-16).

LinearProject maps the 85-d SMPL vector (72 pose + 10 shape + 3 camera) to one 768-d context token; DummyModel is the
identity used by the inference facade to pass pre-computed style embeddings through."""
import torch
from torch import nn


class LinearProject(nn.Module):
    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.model = nn.Linear(input_dim, output_dim)

    def forward(self, x):
        if x.is_cuda:
            from upgpt_b200 import ops
            return ops.linear_small_m(x, self.model.weight, self.model.bias)
        raise RuntimeError("upgpt_b200: LinearProject runs on the CUDA extension only (no CPU fallback)")


class DummyModel(nn.Module):
    def __init__(self, *args, **kwargs):
        # the inference facade swaps style_cond's target to this class but keeps its params ({'device': ...},
        # generate_utils.py:139-142), so the constructor must swallow them like the reference's (poses.py:11-13)
        super().__init__()

    def forward(self, x):
        return x
