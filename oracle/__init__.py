"""CPU oracle (test infrastructure only -- see vcf_oracle.py header)."""
