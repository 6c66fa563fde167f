"""TEST INFRASTRUCTURE (oracle/): stage the reference for the GPU box.

    python oracle/build_ref.py        # needs /root/reference (build container); writes oracle/_ref/reference/

The reference is pure Python on this path (its two CUDA extensions are replaced by stubs on CPU and by the drop-in
mirrors on the GPU), so "building" it is staging its importable files: every *.py and configs/**/*.yaml, copied
unmodified from the read-only checkout into oracle/_ref/ — a git-ignored directory (nothing enters the history) that
is NOT gpurun-ignored, so it travels to the GPU box like the in-tree .so does.  There the parity tests run the
reference's own InfinityGanGenerator / close-loop manager over `spgan_b200.dropin`, and `bench.py --impl reference`
times the reference's own CPU path.  __graft_entry__.build() calls this whenever /root/reference is present.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SPGAN_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref", "reference")
KEEP = (".py", ".yaml", ".yml")


def build(verbose=False):
    if not os.path.isdir(os.path.join(SRC, "models")):
        return None
    n = 0
    for dp, dns, fns in os.walk(SRC):
        dns[:] = [d for d in dns if d not in (".git", "__pycache__")]
        for fn in fns:
            if not fn.endswith(KEEP):
                continue
            src = os.path.join(dp, fn)
            dst = os.path.join(DST, os.path.relpath(src, SRC))
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src) or os.path.getsize(dst) != os.path.getsize(src):
                shutil.copyfile(src, dst)
            n += 1
    if verbose:
        print("staged %d reference files under %s" % (n, DST))
    return DST


if __name__ == "__main__":
    if build(verbose=True) is None:
        sys.exit("reference checkout not found at %s" % SRC)
