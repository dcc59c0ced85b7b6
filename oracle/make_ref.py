#!/usr/bin/env python
"""Recipe for oracle/_ref/: the reference's OWN files for the rerank path, copied unmodified from the read-only
checkout so that they travel to the GPU box (oracle/_ref/ is git-ignored, not gpurun-ignored; nothing from
/root/reference is ever committed).

    python oracle/make_ref.py            # needs /root/reference (the build container); no-op on the GPU box

What is copied and why:
  utilities/diml.py, evaluation/metrics.py   the arithmetic of the path: `bench.py --impl reference` times these very
                                             functions on the host cores (cpu_baseline.kind = "reference") and
                                             tests/ check the restatement (oracle/rerank_oracle.py) against them live
  test_diml_cvt.py, parameters.py            the reference's caller, run UNCHANGED on the B200 through the drop-in
  criteria/, batchminer/, datasampler/       packages (tests/test_gpu_fullpass.py::test_reference_caller_unchanged_on_gpu);
                                             the last three are only imported by the caller's preamble
The reference is pure Python: there is nothing to compile.
"""
import os
import shutil
import sys

REF = os.environ.get("VITRERANK_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "reference")
FILES = ["utilities/diml.py", "evaluation/metrics.py", "test_diml_cvt.py", "parameters.py"]
DIRS = ["criteria", "batchminer", "datasampler"]


def main() -> int:
    if not os.path.isdir(REF):
        print(f"make_ref: {REF} not present (GPU box?) -- keeping whatever oracle/_ref already holds")
        return 0
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    for f in FILES:
        dst = os.path.join(OUT, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, f), dst)
    for d in DIRS:
        shutil.copytree(os.path.join(REF, d), os.path.join(OUT, d),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*~"))
    n = sum(len(fs) for _, _, fs in os.walk(OUT))
    print(f"make_ref: {n} files -> {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
