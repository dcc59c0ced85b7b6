"""Placeholder for the reference's utilities/logger.py (CSV + matplotlib training loggers):
test_diml_cvt.py imports the module (:49) but never uses it on the evaluation path."""
import csv


class CSV_Writer:
    """utilities/logger.py:8-27: append rows to <save_path>_<group>.csv, header once per group."""

    def __init__(self, save_path):
        self.save_path = save_path
        self.written = []
        self.n_written_lines = {}

    def log(self, group, segments, content):
        self.n_written_lines.setdefault(group, 0)
        with open(f"{self.save_path}_{group}.csv", "a") as fh:
            w = csv.writer(fh, delimiter=",")
            if group not in self.written:
                w.writerow(segments)
            for line in content:
                w.writerow(line)
                self.n_written_lines[group] += 1
        self.written.append(group)
