"""Shadow of information_theory.distortion (src/2D-DCT.py:25, :574)."""
from oracle import vcf_oracle as _o
RMSE = _o.rmse
