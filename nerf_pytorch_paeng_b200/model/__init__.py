from .NeRF import NeRF, NeRFModule
from .PositionalEncoding import PositionalEncoding, get_positional_encoder

__all__ = ['NeRF', 'NeRFModule', 'PositionalEncoding', 'get_positional_encoder']
