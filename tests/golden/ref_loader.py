"""Load the Python-2 reference (bifurcation/fourq, impl/*.py) under CPython 3 WITHOUT copying it.

The reference sources are read from /root/reference at run time, rewritten IN MEMORY by the six
purely syntactic Py2->Py3 edits listed in SURVEY.md section 8c, and exec'd into fresh module
objects.  Nothing from the reference is written into this repository; only the *outputs* of its
functions (golden vectors) are committed, by gen_golden.py.

This file is test infrastructure.  It only works where /root/reference exists (the build container);
it is never imported by the product (fourq_b200/) nor by the gpu tests / bench / smoke.

Every edit asserts the number of sites it expects to touch, so a changed reference fails loudly.
"""
import io
import os
import re
import sys
import types
import contextlib

REF_IMPL = os.environ.get("FOURQ_REFERENCE", "/root/reference/impl")


def _sub(src, pattern, repl, expect, flags=0, what=""):
    out, n = re.subn(pattern, repl, src, flags=flags)
    if expect is not None and n != expect:
        raise RuntimeError("ref_loader: edit %r touched %d sites, expected %d" % (what or pattern, n, expect))
    return out


def _py3_common(src):
    # (1) print statements -> print() calls (statement form only: `print "x".format(..)` / bare `print`)
    src = re.sub(r"^(\s*)print[ \t]+(.+)$", r"\1print(\2)", src, flags=re.M)
    src = re.sub(r"^(\s*)print[ \t]*$", r"\1print()", src, flags=re.M)
    # (2) long-literal suffix
    src = re.sub(r"\b(0x[0-9a-fA-F]+|\d+)L\b", r"\1", src)
    return src


def _load(name, src, extra_globals=None):
    mod = types.ModuleType(name)
    mod.__file__ = os.path.join(REF_IMPL, name + ".py") + " (py3-rewritten in memory)"
    if extra_globals:
        mod.__dict__.update(extra_globals)
    sys.modules[name] = mod
    exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    return mod


def load_reference():
    """Returns (fields, curve4q, curve25519, test) modules of the reference, runnable under py3."""
    def read(fn):
        with open(os.path.join(REF_IMPL, fn)) as f:
            return f.read()

    saved = {k: sys.modules.get(k) for k in ("fields", "curve4q", "curve25519", "test")}
    try:
        # ---- fields.py: parses as-is; still pass through the common edits (no-ops)
        fields = _load("fields", _py3_common(read("fields.py")))

        # ---- test.py: print statements only (imports fields)
        test = _load("test", _py3_common(read("test.py")))

        # ---- curve4q.py
        s = _py3_common(read("curve4q.py"))
        # (3) range(n) used as a mutable list
        s = _sub(s, r"^(\s*)(T|d|m|coeff) = range\((\w+)\)$", r"\1\2 = list(range(\3))", 6, re.M, "range->list")
        # (4) integer division (Py2 `/` on ints)
        s = _sub(s, r"reduced = \(reduced - d\[i\]\) / 16", "reduced = (reduced - d[i]) // 16", 1)
        s = _sub(s, r"ind = \[\(abs\(di\) - 1\) / 2 for di in d\]", "ind = [(abs(di) - 1) // 2 for di in d]", 1)
        s = _sub(s, r"sgn = \[di / abs\(di\) for di in d\]", "sgn = [di // abs(di) for di in d]", 1)
        s = _sub(s, r"sgn = \[\(s \+ 1\) / 2 for s in sgn\]", "sgn = [(s + 1) // 2 for s in sgn]", 1)
        # (5) hex codecs
        s = _sub(s, r'str\(encTest\)\.encode\("hex"\)', "bytes(encTest).hex()", 1)
        s = _sub(s, r'bytearray\(Genc\.decode\("hex"\)\)', "bytearray(bytes.fromhex(Genc))", 1)
        curve4q = _load("curve4q", s)

        # ---- curve25519.py
        s = _py3_common(read("curve25519.py"))
        s = _sub(s, r"range\(\(bits\+7\)/8\)", "range((bits+7)//8)", 2)                      # (4)
        s = _sub(s, r"\[ord\(b\) for b in (k|u)\]", r"[b for b in bytearray(\1)]", 2)         # (6)
        s = _sub(s, r"''\.join\(\[chr\(", "bytes(bytearray([(", 1)                             # (6)
        s = _sub(s, r"for i in range\(\(bits\+7\)//8\)\]\)\n\n#{5,} Point Mult",
                 "for i in range((bits+7)//8)]))\n\n########## Point Mult", 1)
        s = _sub(s, r"'([0-9a-f]{64})'\.decode\('hex'\)", r"bytes.fromhex('\1')", None)        # (5)
        curve25519 = _load("curve25519", s)
        return fields, curve4q, curve25519, test
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def run_reference_selftests(mods=None):
    """Runs every self-test of the reference's __main__ blocks; returns the list of printed lines."""
    fields, curve4q, curve25519, test = mods or load_reference()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        for m in (fields, curve4q, curve25519):
            m.test = test
        fields.test_GFp(); fields.test_GFp2(); fields.test_GFp25519()
        curve4q.test_definitions(); curve4q.test_encode(); curve4q.test_reps(); curve4q.test_core()
        curve4q.test_mul_windowed(); curve4q.test_endo(); curve4q.test_recoding()
        curve4q.test_mul_endo(); curve4q.test_dh()
        curve25519.test_x25519(); curve25519.test_dh()
    return [l for l in buf.getvalue().splitlines() if l.strip()]


if __name__ == "__main__":
    lines = run_reference_selftests()
    print("\n".join(lines))
    bad = [l for l in lines if not l.startswith("[PASS]")]
    print("%d lines, %d not PASS" % (len(lines), len(bad)))
    sys.exit(1 if bad else 0)
