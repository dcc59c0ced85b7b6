"""B200 drop-in for the reference's `evaluation` package.  Unlike the reference's
evaluation/__init__.py (which imports faiss and matplotlib at package import, :1) this package
imports nothing heavy, so `from evaluation.metrics import get_metrics_rank` works."""
