"""B200 drop-in for the reference's `utilities` package (only what the DIML rerank path and its
caller test_diml_cvt.py import: diml, misc, logger)."""
