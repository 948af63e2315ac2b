"""ORACLE helper: import the REAL reference modules in-process (this container only).

/root/reference is read-only and does not exist on the GPU box, so this is used
solely by oracle/make_golden.py and by the CPU tests that pin the restatement
(they skip when the tree is absent).  ``munkres`` (PyPI, un-vendored, absent here)
is satisfied by injecting oracle/munkres_ref.py under that module name BEFORE
``rtpe.third_party.group`` is imported (group.py:14).
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("RTPE_REF", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "rtpe", "third_party", "group.py"))


def load_reference():
    """-> (group module, pose_higher_hrnet module) of the unmodified reference."""
    if not reference_available():
        raise FileNotFoundError(REF_ROOT)
    sys.dont_write_bytecode = True
    from . import munkres_ref
    if "munkres" not in sys.modules:
        shim = types.ModuleType("munkres")
        shim.Munkres = munkres_ref.Munkres
        sys.modules["munkres"] = shim
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import rtpe.third_party.group as ref_group
        import rtpe.third_party.pose_higher_hrnet as ref_model
    return ref_group, ref_model


def load_reference_students():
    """-> rtpe.students of the unmodified reference (students.py:595 AttentionStudent)."""
    load_reference()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import rtpe.students as ref_students
    return ref_students
