"""ORACLE helper: import the REAL reference modules in-process.

Search order: ``$RTPE_REF``, ``/root/reference`` (this container; read-only, absent on the GPU
box), ``baseline/_ref`` (a git-ignored copy of the reference's ``rtpe`` package that
``__graft_entry__.build()`` makes when /root/reference is present, so that the reference arm
of ``bench.py`` can time the UNMODIFIED reference on the GPU box's host cores).  Used by
oracle/make_golden*.py, by the CPU tests that pin the restatement (they skip when no tree is
found) and by ``bench.py``'s CPU legs.  ``munkres`` (PyPI, un-vendored, absent here) is
satisfied by injecting oracle/munkres_ref.py under that module name BEFORE
``rtpe.third_party.group`` is imported (group.py:14).
"""
from __future__ import annotations

import os
import sys
import types

_HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTALLED_REF = os.path.join(_HERE, "baseline", "_ref")


def _has_reference(root) -> bool:
    return bool(root) and os.path.isfile(os.path.join(root, "rtpe", "third_party", "group.py"))


def _find_root():
    for cand in (os.environ.get("RTPE_REF"), "/root/reference", INSTALLED_REF):
        if _has_reference(cand):
            return cand
    return os.environ.get("RTPE_REF", "/root/reference")


REF_ROOT = _find_root()


def reference_available() -> bool:
    return _has_reference(REF_ROOT)


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
