"""Multi-resolution hash / dense grid encoder module — same constructor, attributes, ``state_dict`` keys
(``embeddings``, ``offsets``) and forward contract as the reference's ``HashEncoder``
(/root/reference/code/hashencoder/hashgrid.py:140-215); the native op behind it is libnicer_b200.so
(nicer_hash_encode_*), loaded from the in-tree build instead of a JIT at import (backend.py:30-42).
"""
import numpy as np
import torch
import torch.nn as nn

from ..ops import GridMeta, hash_encode


def level_offsets(num_levels, base_resolution, per_level_scale, log2_hashmap_size, input_dim=3):
    """int32 [L+1] prefix sums of min(2^logmap, ceil(base * s^i)^D) (hashgrid.py:160-173)."""
    cap, offs, off = 2 ** log2_hashmap_size, [], 0
    for i in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** i))
        offs.append(off)
        off += min(cap, res ** input_dim)
    offs.append(off)
    return np.array(offs, dtype=np.int32)


class HashEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None):
        super().__init__()
        if desired_resolution is not None:  # overrides per_level_scale (hashgrid.py:145-146)
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim, self.num_levels, self.level_dim = input_dim, num_levels, level_dim
        self.per_level_scale, self.log2_hashmap_size, self.base_resolution = per_level_scale, log2_hashmap_size, base_resolution
        self.output_dim = num_levels * level_dim
        self.max_params = 2 ** log2_hashmap_size
        offsets = level_offsets(num_levels, base_resolution, per_level_scale, log2_hashmap_size, input_dim)
        self.register_buffer("offsets", torch.from_numpy(offsets))
        self.n_params = int(offsets[-1]) * level_dim
        self.embeddings = nn.Parameter(torch.empty(int(offsets[-1]), level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        self.embeddings.data.uniform_(-1e-4, 1e-4)

    def __repr__(self):
        return (f"HashEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"base_resolution={self.base_resolution} per_level_scale={self.per_level_scale} "
                f"params={tuple(self.embeddings.shape)}")

    def grid_meta(self, divide_factor=1.0):
        """Geometry handed to the fused kernels."""
        return GridMeta(L=self.num_levels, C=self.level_dim, H=int(self.base_resolution),
                        S=float(np.log2(self.per_level_scale)), divide_factor=float(divide_factor))

    def forward(self, inputs, size=1):
        """inputs [..., D] in [-size, size] -> [..., L*C]."""
        inputs = (inputs + size) / (2 * size)
        lead = list(inputs.shape[:-1])
        flat = inputs.view(-1, self.input_dim)
        out = hash_encode(flat, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                          flat.requires_grad)
        return out.view(lead + [self.output_dim])
