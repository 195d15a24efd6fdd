"""NeRF positional encoding [x, sin(2^k x), cos(2^k x)]_k (reference: model/embedder.py:5-37,71-88).
Only used on the un-fused fallback path; the fused kernels generate the encoding in registers."""
import torch


class Embedder:
    def __init__(self, input_dims, num_freqs, include_input=True):
        self.freqs = [2.0 ** k for k in range(num_freqs)]
        self.include_input = include_input
        self.out_dim = input_dims * (2 * num_freqs + (1 if include_input else 0))

    def embed(self, x):
        parts = [x] if self.include_input else []
        for f in self.freqs:
            parts += [torch.sin(x * f), torch.cos(x * f)]
        return torch.cat(parts, -1)


def get_embedder(multires, input_dims=3, embed_type="nerf"):
    if embed_type != "nerf":
        raise NotImplementedError(f"embedding_method={embed_type!r}: only 'nerf' is on the hot path (no shipped conf uses another)")
    e = Embedder(input_dims, multires)
    return e.embed, e.out_dim
