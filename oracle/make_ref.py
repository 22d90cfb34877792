"""ORACLE recipe (test infrastructure): stage the UPSTREAM hot-path sources, untouched, into the
git-ignored ``oracle/_ref/`` so that they travel to the GPU box with the snapshot (ignored files ship,
``.git`` history stays source-only).

    python oracle/make_ref.py          # needs the upstream checkout at /root/reference

Called from ``__graft_entry__.build()`` whenever ``/root/reference`` is present (the build container);
on the GPU box only the staged copy exists.  Nothing is edited: the files are byte copies (checked by
sha256 into ``oracle/_ref/MANIFEST.json``), read through ``oracle/upstream.py`` by ``tests/`` and by
bench.py's reference / baseline legs -- never by the product package.

Staged: models/{__init__,module,Effi_MVS_plus,update,loss}.py, utils.py
(module.py:6-7 imports it), misc/fusion.py.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["models/__init__.py", "models/module.py", "models/Effi_MVS_plus.py", "models/update.py", "models/loss.py",
         "utils.py", "misc/fusion.py"]


def stage(upstream: str = "/root/reference", dest: str = DEST) -> bool:
    if not os.path.isdir(os.path.join(upstream, "models")):
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(upstream, rel), os.path.join(dest, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        os.chmod(dst, 0o644)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"upstream": "bdwsq1996/Effi-MVS-plus", "files": manifest}, f, indent=1)
    return True


if __name__ == "__main__":
    ok = stage(*(sys.argv[1:2]))
    print("oracle/_ref staged" if ok else "upstream checkout not found; oracle/_ref left as it is")
