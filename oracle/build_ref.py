"""Stages the reference's own implementation of the path under ``oracle/_ref/`` (test / baseline infrastructure, never
imported by the product).

The reference is Python: "building" it means copying the one file the path lives in,
``framework/domain_adaptation/methods/prototype_handler.py`` (needs only torch), from where it lies under
``/root/reference`` into ``oracle/_ref/`` -- a git-ignored directory that travels to the GPU box like a built ``.so``
does.  ``bench.py --impl reference`` and the ``cpu_baseline`` leg then time the REAL class (``kind: "reference"``)
instead of the oracle port.  No reference source is committed to this repository.
"""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("ONDA_REFERENCE", "/root/reference")
FILES = ["framework/domain_adaptation/methods/prototype_handler.py"]
OUT = os.path.join(HERE, "_ref")


def build_ref() -> bool:
    """Copy the files if the reference tree is present (authoring container); keep what is there otherwise."""
    if not os.path.isdir(REF_ROOT):
        return os.path.exists(os.path.join(OUT, "prototype_handler.py"))
    os.makedirs(OUT, exist_ok=True)
    for rel in FILES:
        shutil.copyfile(os.path.join(REF_ROOT, rel), os.path.join(OUT, os.path.basename(rel)))
    return True


def load_reference_handler():
    """The reference's ``prototype_handler`` class from ``oracle/_ref`` or None if it was never staged."""
    path = os.path.join(OUT, "prototype_handler.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("onda_reference_prototype_handler", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.prototype_handler


if __name__ == "__main__":
    print("staged" if build_ref() else "reference tree absent and nothing staged")
