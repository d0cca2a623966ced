"""Developer scripts' access to the tensor-core IPA test inputs and truth function, which live under tests/ (pinned to the oracle
there): tests/ipa_tc_reference.py."""
import os, sys
_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _root)
sys.path.insert(0, os.path.join(_root, "tests"))
import torch  # noqa: E402
from se3diff_b200 import ops, _lib as L  # noqa: E402,F401
from ipa_tc_reference import H, dk, D, make, head_major, ref, split  # noqa: E402,F401

dev = "cuda"
