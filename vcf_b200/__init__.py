"""vcf_b200 -- B200-native colour + block-DCT + deadzone path of VCF.

Public API: :mod:`vcf_b200.codec` (arrays in, arrays out) and the drop-in
spatial-transform module ``vcf_b200/plugin/2D-DCT-B200.py`` for the reference's
CoDec chain.  The arithmetic lives in ``libvcfb200.so`` (CUDA, sm_100a)."""
from .codec import (  # noqa: F401
    Codec, ColorCodec, decode_frames, encode_frames, padded_dims, perceptual_weights, pinned_empty, rd_stats,
)
from ._lib import VcfbError  # noqa: F401

__version__ = "0.2.0"
