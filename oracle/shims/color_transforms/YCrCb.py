"""Shadow of color_transforms.YCrCb (src/YCrCb.py:11-12): 8-bit OpenCV path."""
from oracle import vcf_oracle as _o
name = "YCrCb"
from_RGB = _o.ycrcb_from_rgb_u8
to_RGB = _o.ycrcb_to_rgb_u8
