"""
TEST INFRASTRUCTURE ONLY.  Packs the UNTOUCHED reference package (`indigo/` and
`examples/pics.py`, whose inline Transform classes are the -O1..-O3 recipe) into
oracle/_ref/reference_pkg.zip, so that the -m gpu tests can run the reference's own
operators.py / transforms.py / test suites on the B200 backend on the GPU box, which has no
/root/reference.  oracle/_ref/ is git-ignored build output (it travels with the gpurun
snapshot like a built .so); no reference source enters the repository's history or tree.

    python oracle/stage_reference.py /root/reference oracle/_ref/reference_pkg.zip
"""
import os
import sys
import zipfile


def stage(reference, out):
    keep = []
    for root, dirs, files in os.walk(os.path.join(reference, "indigo")):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".h")):
                keep.append(os.path.join(root, f))
    keep.append(os.path.join(reference, "examples", "pics.py"))
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    with zipfile.ZipFile(out, "w", zipfile.ZIP_DEFLATED) as z:
        for path in sorted(keep):
            info = zipfile.ZipInfo(os.path.relpath(path, reference), date_time=(2020, 1, 1, 0, 0, 0))
            info.compress_type = zipfile.ZIP_DEFLATED
            with open(path, "rb") as fh:
                z.writestr(info, fh.read())
    return len(keep)


if __name__ == "__main__":
    n = stage(sys.argv[1], sys.argv[2])
    print("staged %d reference files in %s" % (n, sys.argv[2]))
