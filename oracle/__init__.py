"""Test-only CPU oracle (see diff3d_oracle.py header).  Not importable from the product package."""
