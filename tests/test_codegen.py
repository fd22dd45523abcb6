"""The committed CUDA codelets must be exactly what the generator prints."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_codelets_in_sync_with_generator():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "vcf_b200", "codegen", "gen_cuda.py"), "--check"], cwd=ROOT)
    assert r.returncode == 0, "run `make codelets` and commit vcf_b200/csrc/dct_codelets.cuh"
