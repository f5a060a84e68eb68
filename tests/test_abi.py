"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every
symbol include/adflow_b200.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

from adflow_b200 import _lib, make_params
from adflow_b200.params import AdfbParams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "adflow_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(adfb_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_exported():
    L = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "libadflow_b200.so does not export %s" % s
    assert set(_lib.ABI_SYMBOLS) == set(syms)


def test_params_struct_size_matches_header():
    # 64 doubles? count from the header: keep the ctypes twin in sync with the C struct
    txt = open(os.path.join(ROOT, "include", "adflow_b200.h")).read()
    body = txt[txt.index("typedef struct AdfbParams {"):txt.index("} AdfbParams;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    nd = 0
    for decl in re.findall(r"double\s+([^;]+);", body):
        for name in decl.split(","):
            m = re.search(r"\[(\d+)\]", name)
            nd += int(m.group(1)) if m else 1
    ni = len(re.findall(r"int32_t\s+\w+\s*;", body))
    assert C.sizeof(AdfbParams) == 8 * nd + 4 * ni


def test_fails_loudly_without_gpu():
    L = _lib.load()
    if L.adfb_device_count() > 0:
        pytest.skip("a GPU is present")
    rc = L.adfb_init(0, None, 0, 1)
    assert rc != 0
    buf = C.create_string_buffer(512)
    L.adfb_last_error(buf, 512)
    assert b"no CUDA device" in buf.value
    prm = make_params()
    assert L.adfb_set_params(C.byref(prm)) != 0  # not initialised -> error, never a CPU path
    assert L.adfb_residual(1, 8 | 16) != 0


def test_integration_doc_binds_every_abi_symbol():
    """INTEGRATION.md shows the ISO_C_BINDING interface a maintainer adds: it has to cover the whole C ABI"""
    import os
    import re

    from adflow_b200._lib import ABI_SYMBOLS

    doc = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "INTEGRATION.md")).read()
    bound = set(re.findall(r'bind\(c,\s*name="(\w+)"\)', doc))
    assert sorted(set(ABI_SYMBOLS) - bound) == []


def test_integration_doc_interface_matches_header_arity():
    """every Fortran interface in INTEGRATION.md has as many dummies as its C prototype has parameters"""
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "adflow_b200.h")).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|long long|double)\s+(adfb_\w+)\s*\(([^;{]*?)\)\s*;", h, re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("void", "") else len(re.split(r",(?![^()]*\))", args))
    doc = re.sub(r"&\s*\n\s*", "", open(os.path.join(root, "INTEGRATION.md")).read())
    fort = {}
    for m in re.finditer(r"function\s+(adfb_\w+)\s*\(([^)]*)\)", doc):
        a = m.group(2).strip()
        fort[m.group(1)] = 0 if not a else len(a.split(","))
    assert len(protos) >= 50
    assert sorted(k for k in protos if k not in fort) == []
    assert [(k, protos[k], fort[k]) for k in protos if protos[k] != fort[k]] == []


def test_struct_layouts_match_the_header_field_by_field():
    """sizeof / offsetof of AdfbParams, AdfbAnkParams and AdfbSubface as gcc sees the header == the ctypes twins the
    Python layer passes, and the field ORDER of the bind(c) types shown in INTEGRATION.md is the header's"""
    import subprocess

    from adflow_b200._lib import AdfbSubface
    from adflow_b200.params import AdfbAnkParams

    work = os.path.join(ROOT, "oracle", "_ref", "_selftest")
    os.makedirs(work, exist_ok=True)
    twins = {"AdfbParams": AdfbParams, "AdfbAnkParams": AdfbAnkParams, "AdfbSubface": AdfbSubface}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "adflow_b200.h"', 'int main(void) {']
    for name, T in twins.items():
        lines.append('printf("%s sizeof %%zu\\n", sizeof(%s));' % (name, name))
        for f, _t in T._fields_:
            cf = {"pad_": "pad_"}.get(f, f)
            lines.append('printf("%s %s %%zu\\n", offsetof(%s, %s));' % (name, f, name, cf))
    lines += ["return 0; }"]
    src = os.path.join(work, "layout.c")
    with open(src, "w") as fh:
        fh.write("\n".join(lines))
    exe = os.path.join(work, "layout")
    subprocess.check_call(["gcc", "-I" + os.path.join(ROOT, "include"), "-o", exe, src])
    out = subprocess.check_output([exe]).decode().split("\n")
    got = {}
    for ln in out:
        if ln.strip():
            a, b, c = ln.split()
            got[(a, b)] = int(c)
    for name, T in twins.items():
        assert got[(name, "sizeof")] == C.sizeof(T), name
        for f, _t in T._fields_:
            assert got[(name, f)] == getattr(T, f).offset, (name, f)
    # Fortran types: same field order as the header
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for name, T in (("AdfbParams", AdfbParams), ("AdfbAnkParams", AdfbAnkParams)):
        body = doc[doc.index("type, bind(c) :: %s" % name):doc.index("end type %s" % name)]
        body = re.sub(r"&\s*\n\s*", "", body)
        names = []
        for decl in re.findall(r"::\s*([^\n]+)", body)[1:]:
            names += [re.sub(r"\(.*\)", "", x).strip() for x in decl.split(",")]
        assert [n.lower() for n in names] == [f.lower() for f, _t in T._fields_], name
