"""Recipe for oracle/_ref/: the UNMODIFIED reference files of the hot path, copied byte for byte from /root/reference so that the GPU
box (which has no /root/reference) can time and check the reference's own classes (bench.py --impl reference -> "kind": "reference";
tests/_refload.py with MHIM_REFERENCE_ROOT=oracle/_ref).  oracle/_ref/ is git-ignored (reference sources never enter the history) but
not gpurun-ignored, so it travels with the snapshot like the built .so.  Test / measurement infrastructure only: nothing under
mhim-mil_b200/ imports it.

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference is present)
"""
import os
import shutil
import sys

SRC = os.environ.get("MHIM_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ["modules/abmil.py", "modules/emb_position.py", "modules/mhim.py", "modules/dsmil.py", "modules/transmil.py",
         "modules/nystrom_attention.py", "modules/mhim_modules/__init__.py", "modules/mhim_modules/baseline.py",
         "modules/mhim_modules/masking.py", "modules/mhim_modules/scoring.py", "modules/mhim_modules/merge.py",
         "modules/mhim_modules/losses.py", "modules/mhim_modules/utils.py", "engines/common_mil.py", "modules/dtfd.py",
         "modules/clam.py", "modules/topk/__init__.py", "modules/topk/svm.py", "modules/topk/functional.py", "modules/topk/logarithm.py",
         "modules/topk/utils.py", "modules/topk/polynomial/__init__.py", "modules/topk/polynomial/sp.py", "modules/topk/polynomial/grad.py",
         "modules/topk/polynomial/multiplication.py", "modules/topk/polynomial/divide_conquer.py"]


def main() -> int:
    if not os.path.isfile(os.path.join(SRC, "modules", "mhim.py")):
        print(f"oracle/make_ref.py: {SRC} not present -- keeping whatever oracle/_ref/ already holds")
        return 0
    n = 0
    for rel in FILES:
        src = os.path.join(SRC, rel)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        n += 1
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as f:
        f.write(f"{n} files copied unmodified from {SRC} by oracle/make_ref.py (DearCaat/MHIM-MIL @ 9d0c91a)\n")
    print(f"oracle/_ref: {n} reference files")
    return 0


if __name__ == "__main__":
    sys.exit(main())
