"""CPU: the C-ABI library loads and exports every symbol include/vscuda.h declares; the host mirror's
argument checks behave like the reference's panics; and the product path fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "vscuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vs_[a-z0-9_A-Z]+)\s*\(", text)))


def test_header_symbols_exported(pkg):
    names = _declared()
    assert len(names) >= 40
    L = ctypes.CDLL(pkg.lib_path())
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(pkg._lib.SYMBOLS) == names


def test_no_torch_types_in_header():
    text = open(os.path.join(ROOT, "include", "vscuda.h")).read()
    assert "torch" not in text.lower() and "at::" not in text and "Tensor" not in text


def test_product_never_imports_oracle():
    pkgdir = os.path.join(ROOT, "go-vectorsearch_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cc", ".cpp", ".go")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.lower(), f"{f} mentions the oracle"


def test_host_mirror_panics_like_reference(pkg):
    c = pkg.compute
    with pytest.raises(c.ComputePanic, match="vector columns are empty"):
        c.NewVector(np.zeros(8, np.uint8))          # compute.go:12-14
    with pytest.raises(c.ComputePanic, match="matrix rows are empty"):
        c.NewMatrix([])                             # compute.go:25-27
    with pytest.raises(c.ComputePanic):
        c.NewMatrix([np.zeros(12, np.uint8), np.zeros(13, np.uint8)])


def test_fails_loudly_without_gpu(pkg):
    """No CPU fallback: on a box without a CUDA device every compute entry point raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.BackendUnavailable):
        pkg.compute.NewMatrix(np.zeros((2, 16), np.uint8))
    with pytest.raises(pkg.BackendUnavailable):
        pkg.compute.QuantizeVectorFloat32(np.ones(4, np.float32))


def _split_top_level(argtext):
    """Split an argument list on commas that are not nested inside parentheses / brackets / braces."""
    parts, depth, cur = [], 0, ""
    for ch in argtext:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur.strip())
    return parts


def _call_args(src, start):
    """Text between the parenthesis opening at src[start] and its match."""
    depth = 0
    for i in range(start, len(src)):
        if src[i] == "(":
            depth += 1
        elif src[i] == ")":
            depth -= 1
            if depth == 0:
                return src[start + 1:i]
    raise AssertionError("unbalanced call")


def test_go_shim_calls_match_header():
    """Go is not in this image, so the cgo shim cannot be compiled here: at least every C.vs_* call in goshim/*.go must
    name a function include/vscuda.h declares and pass as many arguments as its prototype has."""
    text = open(os.path.join(ROOT, "include", "vscuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\bVS_API\s+[\w\s\*]+?\b(vs_\w+)\s*\(", text):
        args = _call_args(text, m.end() - 1).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(_split_top_level(args))
    shim_dir = os.path.join(ROOT, "go-vectorsearch_b200", "goshim")
    calls = 0
    for f in sorted(os.listdir(shim_dir)):
        if not f.endswith(".go"):
            continue
        src = open(os.path.join(shim_dir, f)).read()
        src = re.sub(r"//[^\n]*", "", src)
        for m in re.finditer(r"\bC\.(vs_\w+)\s*\(", src):
            name = m.group(1)
            assert name in protos, f"{f}: C.{name} is not declared in vscuda.h"
            n = len(_split_top_level(_call_args(src, m.end() - 1)))
            assert n == protos[name], f"{f}: C.{name} called with {n} arguments, prototype has {protos[name]}"
            calls += 1
    assert calls >= 20
