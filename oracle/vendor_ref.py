"""TEST INFRASTRUCTURE / CPU BASELINE ONLY -- recipe that makes the UNMODIFIED reference available to bench.py's
reference arm on the GPU box.

    python -m oracle.vendor_ref            # in the build container, where /root/reference is mounted

Copies the reference's pure-Python package ``machin/`` (0.9 MB, no build step) byte for byte from where it lies under
``/root/reference`` into ``oracle/_ref/machin``.  ``oracle/_ref/`` is git-ignored (reference sources never enter the
repository's history) but not gpurun-ignored, so it travels to the GPU box with the snapshot like a built ``.so``.
``__graft_entry__.build()`` runs this when the reference is mounted.  Only ``bench.py --impl reference`` /
``cpu_baseline`` (through oracle/ref_learner.py) and the golden-vector generator import the result; nothing under
``machin_b200/`` does.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MACHIN_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def vendor(quiet=False) -> bool:
    src = os.path.join(SRC, "machin")
    if not os.path.isdir(src):
        if not quiet:
            print(f"reference not mounted at {SRC}: nothing to do (the GPU box uses the prebuilt oracle/_ref)")
        return os.path.isdir(os.path.join(DST, "machin"))
    dst = os.path.join(DST, "machin")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Unmodified copy of iffiX/machin's `machin/` package made by oracle/vendor_ref.py; git-ignored build output.\n")
    if not quiet:
        print(f"vendored {src} -> {dst}")
    return True


if __name__ == "__main__":
    sys.exit(0 if vendor() else 1)
