"""Puts the repository root on sys.path so that the drop-in modules beside this file (imported with cwd = image/, exactly
as the reference's train.py / generate.py import theirs) can reach the ``reed_b200`` package without an install."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
