"""TEST / BASELINE INFRASTRUCTURE ONLY -- packs the UNMODIFIED reference's own Python modules of the pairwise-order path
into ``oracle/_ref/instaorder_ref.zip`` so that the reference itself (not a port) can be timed on the GPU box's host
cores by ``bench.py --impl reference`` / ``cpu_baseline`` (``kind: "reference"``).

The reference is pure Python: there is nothing to compile, "building" it is archiving the modules
where they lie under ``/root/reference`` -- ``inference.py``, ``utils/``, ``models/``, ``midas/`` (the packages its
``import utils / inference / models`` chain pulls in) -- byte for byte.  The archive is an artefact like a compiled
``.so``: ``oracle/_ref/`` is git-ignored (no reference source enters the history) but not gpurun-ignored, so it travels
to the GPU box, where ``/root/reference`` does not exist.  ``oracle/ref_shim.py`` imports from the archive through
``zipimport`` when the tree is absent.  Run by ``__graft_entry__.build()`` in the build container; never imported by
anything under ``instaorder_b200/``."""
import os
import sys
import zipfile

SRC = os.environ.get("INSTAORDER_REFERENCE_SRC", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "instaorder_ref.zip")
TOP = ("inference.py",)
PACKAGES = ("utils", "models", "midas")


def build(src=SRC, out=OUT):
    """Returns the archive path, or None when the reference tree is not present (GPU box: the shipped file is used)."""
    if not os.path.isfile(os.path.join(src, "inference.py")):
        return None
    files = [t for t in TOP]
    for pkg in PACKAGES:
        for root, _, names in os.walk(os.path.join(src, pkg)):
            for n in sorted(names):
                if n.endswith(".py"):
                    files.append(os.path.relpath(os.path.join(root, n), src))
    # ``midas/`` has no __init__.py (an implicit namespace package on disk); zipimport needs a regular package, so an
    # EMPTY __init__.py is added for such directories -- the only bytes in the archive that are not the reference's
    synth = [os.path.join(d, "__init__.py") for d in sorted({os.path.dirname(f) for f in files if os.path.dirname(f)})
             if os.path.join(d, "__init__.py") not in files]
    os.makedirs(os.path.dirname(out), exist_ok=True)
    tmp = out + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for rel in sorted(files):
            zi = zipfile.ZipInfo(rel, date_time=(2020, 1, 1, 0, 0, 0))      # reproducible archive
            zi.compress_type = zipfile.ZIP_DEFLATED
            with open(os.path.join(src, rel), "rb") as f:
                z.writestr(zi, f.read())
        for rel in synth:
            z.writestr(zipfile.ZipInfo(rel, date_time=(2020, 1, 1, 0, 0, 0)), b"")
    os.replace(tmp, out)
    return out


if __name__ == "__main__":
    p = build()
    print("wrote %s (%d bytes)" % (p, os.path.getsize(p)) if p else "reference tree not found at %s" % SRC)
    sys.exit(0)
