import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _oracle_path  # noqa: F401  (puts the repo root on sys.path)
