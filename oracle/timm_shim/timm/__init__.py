"""Minimal stand-in for the un-vendored, unpinned `timm` dependency of the reference.

TEST INFRASTRUCTURE ONLY.  The reference imports three classes from timm
(/root/reference/image/models/sit.py:13); timm is not installed in this image and
cannot be installed (no network).  This shim restates their published semantics
(timm 0.9.x-1.0.x) so that the reference's sit.py imports UNMODIFIED when
oracle/make_golden.py generates the committed fixtures under tests/golden/.
"""
