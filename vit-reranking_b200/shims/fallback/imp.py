"""`imp` was removed in Python 3.12; test_diml_cvt.py:8 imports it and never uses it."""
