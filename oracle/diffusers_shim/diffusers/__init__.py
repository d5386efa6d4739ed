"""Minimal stand-in for the un-vendored `diffusers==0.35.2` dependency.

TEST INFRASTRUCTURE ONLY.  It exists so that the reference file
`/root/reference/chronoedit_diffusers/transformer_chronoedit.py` can be
imported and executed UNMODIFIED on CPU when generating golden vectors
(`tests/golden/make_golden.py`).  It provides exactly the symbols that file
imports (transformer_chronoedit.py:23-32); every class restates the published
diffusers 0.35.2 semantics ("[diffusers-mem]").
Nothing in the product package imports this.
"""
__version__ = "0.35.2+shim"
