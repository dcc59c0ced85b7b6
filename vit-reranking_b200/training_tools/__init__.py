"""B200 drop-in for the rerank core of the reference's `training_tools` package (val.py)."""
