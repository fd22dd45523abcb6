"""Shadow of color_transforms.YCoCg (src/2D-DCT.py:22-23, src/YCoCg.py:11-12)."""
from oracle import vcf_oracle as _o
name = "YCoCg"
from_RGB = _o.ycocg_from_rgb
to_RGB = _o.ycocg_to_rgb
