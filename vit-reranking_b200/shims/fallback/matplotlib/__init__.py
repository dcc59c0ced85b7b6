"""Minimal stand-in used only when matplotlib is not installed: test_diml_cvt.py:10-12 calls
matplotlib.use('agg') and imports pyplot, nothing else on the evaluation path."""


def use(*args, **kwargs):
    return None
