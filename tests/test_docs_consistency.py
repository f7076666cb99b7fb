"""The binding shown in INTEGRATION.md and the header must name the same pd_config fields, in the same
order, as the binding the package uses (a drifted stub would silently corrupt every parameter)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _package_fields():
    from pdabm_b200 import _lib
    return [n for n, _ in _lib.pd_config._fields_]


def test_integration_md_stub_matches_binding():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = text[text.index("class pd_config(C.Structure)"):text.index("lib = C.CDLL")]
    assert re.findall(r'\("(\w+)",\s*C\.', block) == _package_fields()


def test_header_struct_matches_binding():
    text = open(os.path.join(ROOT, "include", "pdabm.h")).read()
    body = text[text.index("typedef struct pd_config {"):text.index("} pd_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"\b(?:u?int(?:8|32|64)_t|float)\s+(\w+)(?:\[\d+\])?\s*;", body)
    assert names == _package_fields()


def test_every_exported_function_is_documented_in_the_header():
    from pdabm_b200 import _lib
    text = open(os.path.join(ROOT, "include", "pdabm.h")).read()
    for name in _lib.EXPORTS:
        assert re.search(r"\b%s\s*\(" % name, text), name
