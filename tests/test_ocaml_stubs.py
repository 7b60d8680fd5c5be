"""The OCaml binding (ocaml/) without an OCaml toolchain.

The image has no `ocaml`, `dune` or caml headers, so the binding cannot be built against the real
runtime.  What can be checked is checked:

* CPU: `ocaml/phylo_b200_stubs.c` compiles with `-Wall -Wextra -Werror` against `include/phylo_b200.h`
  and a mock of the runtime's C interface (`tests/ocaml_mock/caml/`), and links against the C-ABI library;
* CPU: every `external` of `ocaml/phylo_b200.ml` names a `CAMLprim` of the stubs with the same number of
  arguments (and a `*_bytecode` twin taking `(value *, int)` above five arguments), and every `CAMLprim`
  is bound by an `external`;
* GPU: `tests/ocaml_mock/stubs_driver.c` plays the OCaml side — it calls the stubs in the order
  `Phylo_b200.pruning_sites` / `Felsenstein(..).evaluate` do, with Bigarray descriptors built in C — and
  checks the reference's RNG-free fixtures, the exception mapping, the runtime-lock pairing and the finaliser.
"""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUBS = os.path.join(ROOT, "ocaml", "phylo_b200_stubs.c")
ML = os.path.join(ROOT, "ocaml", "phylo_b200.ml")
MOCK = os.path.join(ROOT, "tests", "ocaml_mock")
EXE = os.path.join(MOCK, "_build", "stubs_driver")


def build_driver():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.run(["/usr/bin/gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror",
                    "-I" + os.path.join(ROOT, "include"), "-I" + MOCK,
                    STUBS, os.path.join(MOCK, "stubs_driver.c"), "-o", EXE,
                    "-L" + os.path.join(ROOT, "phylogenetics_b200"), "-lphylo_b200", "-lm",
                    "-Wl,-rpath,$ORIGIN/../../../phylogenetics_b200"], check=True)
    return EXE


def test_stubs_compile_against_the_header_and_link():
    build_driver()
    assert os.path.exists(EXE)


def _primitives():
    """name -> number of `value` parameters (or 'bytecode') of every CAMLprim in the stubs."""
    src = open(STUBS).read()
    prims = {}
    for name, params in re.findall(r"CAMLprim\s+value\s+(\w+)\s*\(([^)]*)\)", src):
        params = [p.strip() for p in params.split(",")]
        if params == ["value *a", "int n"]:
            prims[name] = "bytecode"
        else:
            assert all(re.fullmatch(r"value \w+", p) for p in params), (name, params)
            prims[name] = len(params)
    return prims


def _externals():
    """OCaml name -> (arity, [C names]) of every `external` in phylo_b200.ml."""
    src = re.sub(r"\(\*.*?\*\)", "", open(ML).read(), flags=re.S)
    out = {}
    for name, ty, names in re.findall(r"external\s+(\w+)\s*:\s*(.*?)=\s*((?:\"\w+\"\s*)+)", src, flags=re.S):
        arity = len([t for t in ty.split("->")]) - 1
        out[name] = (arity, re.findall(r"\"(\w+)\"", names))
    return out


def test_every_external_has_a_stub_of_the_same_arity():
    prims, exts = _primitives(), _externals()
    assert len(exts) >= 18 and len(prims) >= 21
    bound = set()
    for name, (arity, cnames) in exts.items():
        if arity > 5:                       # OCaml's rule: byte-code twin first, native second
            assert len(cnames) == 2, name
            assert prims.get(cnames[0]) == "bytecode", (name, cnames[0])
            assert prims.get(cnames[1]) == arity, (name, cnames[1], prims.get(cnames[1]), arity)
        else:
            assert len(cnames) == 1, name
            assert prims.get(cnames[0]) == arity, (name, cnames[0], prims.get(cnames[0]), arity)
        bound.update(cnames)
    assert bound == set(prims), sorted(set(prims) ^ bound)


def test_driver_fails_loudly_without_a_gpu():
    """No CPU fallback behind the stubs either: pb_stub_create raises Failure when there is no device."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    build_driver()
    out = subprocess.run([EXE], capture_output=True, text=True)
    assert out.returncode == 2 and "no CUDA device" in out.stderr, out.stdout + out.stderr


@pytest.mark.gpu
def test_stubs_run_on_the_gpu_against_the_mock_runtime():
    build_driver()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all ok" in out.stdout and "FAIL" not in out.stdout, out.stdout
