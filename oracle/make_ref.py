"""Recipe for `oracle/_ref/` (git-ignored; it travels to the GPU box with the repository snapshot like a built .so):
a VERBATIM copy of exactly those reference files that importing the GET model needs, taken from the read-only tree
/root/reference. Run in the build container:   python oracle/make_ref.py

The file list is not hand-written: the reference is imported once (oracle/ref_import.py) and every module it loaded from
the reference tree is copied, plus losses.py (the loss the fitter uses). A MANIFEST.json with the sha256 of every copied
file is written next to them. No reference source is committed to the repository."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
EXTRA = ["losses.py"]


def main() -> int:
    if not os.path.isdir(os.path.join(SRC, "Models")):
        print("make_ref: %s not present (only available in the build container); nothing to do" % SRC)
        return 0
    os.environ["GET_REFERENCE_ROOT"] = SRC
    from oracle import ref_import
    ref_import.import_reference()
    files = set(EXTRA)
    for m in list(sys.modules.values()):
        f = getattr(m, "__file__", None)
        if f and os.path.abspath(f).startswith(SRC + os.sep):
            files.add(os.path.relpath(os.path.abspath(f), SRC))
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for rel in sorted(files):
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(src, "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    print("make_ref: %d files copied to %s" % (len(manifest), DST))
    return 0


if __name__ == "__main__":
    sys.exit(main())
