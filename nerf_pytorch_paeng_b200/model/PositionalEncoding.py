"""Drop-in for the reference's model/PositionalEncoding.py (file:line refs are into the reference).

``get_positional_encoder(L) -> (callable, out_dim)`` as PositionalEncoding.py:33-36; the callable
maps a CUDA fp32 tensor [..., 3] to [..., 3+6L] = [x, sin(2^k x), cos(2^k x)]_{k<L} (:18-30) with
one launch of the nb_posenc kernel.  Inside render_rays the encoding is never materialised for
the tensor-core path (it is generated in the MLP's operand producer).
"""
from ..engine import get_engine


class PositionalEncoding:
    def __init__(self, L: int):
        self.L = int(L)
        self.input_dims = 3
        self.out_dim = 3 + 6 * self.L          # PositionalEncoding.py:12-24

    def embed(self, inputs):
        return get_engine(inputs.device).posenc(inputs, self.L)


def get_positional_encoder(L: int):
    embedder_obj = PositionalEncoding(L)

    def pos_encoder(x, eo=embedder_obj):
        return eo.embed(x)

    pos_encoder.L = embedder_obj.L             # lets render_rays recover L for the fused path
    return pos_encoder, embedder_obj.out_dim
