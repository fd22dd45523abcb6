"""Shadow of scalar_quantization.deadzone_quantization (src/deadzone.py:10-11)."""
from oracle import vcf_oracle as _o
name = "deadzone"
Deadzone_Quantizer = _o.DeadzoneQuantizer
