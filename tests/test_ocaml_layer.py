"""The OCaml stub layer (ocaml/) cannot be linked here -- there is no OCaml toolchain --
so what can be checked on a CPU box is checked: the C stubs compile as C against
stand-ins for the caml headers, every library call they make is declared in
include/soundml_b200.h with that arity, and every ``external`` of soundml_b200.ml names a
CAMLprim of the stub file with the matching number of arguments (and a bytecode twin
where OCaml requires one: more than five arguments, resample_stubs.c:410-422)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUBS = os.path.join(ROOT, "ocaml", "soundml_b200_stubs.c")
ML = os.path.join(ROOT, "ocaml", "soundml_b200.ml")


def test_stubs_compile_against_the_caml_shim():
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not found")
    res = subprocess.run(
        [gcc, "-std=c11", "-fsyntax-only", "-Wall", "-Wextra", "-Werror",
         "-I", os.path.join(ROOT, "oracle", "caml_shim"), "-I", os.path.join(ROOT, "include"), STUBS],
        capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def _prims():
    src = open(STUBS).read()
    out = {}
    for m in re.finditer(r"CAMLprim value (\w+)\(([^)]*)\)", src):
        args = [a for a in m.group(2).split(",") if a.strip()]
        out[m.group(1)] = None if "argv" in m.group(2) else len(args)
    return out


def test_every_external_names_a_stub_of_the_right_arity():
    prims = _prims()
    src = open(ML).read()
    seen = 0
    for m in re.finditer(r"external\s+\w+\s*:(.*?)=\s*((?:\"\w+\"\s*)+)", src, re.S):
        names = re.findall(r"\"(\w+)\"", m.group(2))
        sig = re.sub(r"\(\*.*?\*\)", "", m.group(1), flags=re.S)
        depth, arrows = 0, 0
        for i, ch in enumerate(sig):               # arrows outside parentheses = arguments
            depth += ch == "("
            depth -= ch == ")"
            if ch == "-" and sig[i:i + 2] == "->" and depth == 0:
                arrows += 1
        native = names[-1]
        assert native in prims, native
        assert prims[native] == arrows, (native, prims[native], arrows)
        if arrows > 5:
            assert len(names) == 2 and names[0] == native + "_bc" and names[0] in prims, names
        else:
            assert len(names) == 1, names
        seen += 1
    assert seen >= 20


def test_stub_calls_match_the_header():
    header = open(os.path.join(ROOT, "include", "soundml_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = {}
    for m in re.finditer(r"\b(smb_\w+)\s*\(([^;{]*?)\)\s*;", header, re.S):
        args = m.group(2).strip()
        declared[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    src = re.sub(r"/\*.*?\*/", "", open(STUBS).read(), flags=re.S)
    calls = 0
    for m in re.finditer(r"\b(smb_\w+)\s*\(", src):
        name = m.group(1)
        if name in ("smb_ml_raise",):
            continue
        assert name in declared, name
        i, depth, args = m.end(), 1, 1
        while depth:
            ch = src[i]
            depth += ch == "("
            depth -= ch == ")"
            args += ch == "," and depth == 1
            i += 1
        if src[m.end():i - 1].strip() == "":
            args = 0
        assert args == declared[name], (name, args, declared[name])
        calls += 1
    assert calls >= 25
