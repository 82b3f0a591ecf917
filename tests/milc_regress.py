"""The reference's regression procedure, restated in Python (test infrastructure).

MILC checks an application by running it on a shipped sample input and comparing selected
output with a trusted sample, field by field, against per-field absolute tolerances
(Make_test_template:86-92, headtail.pl, diffn3.pl).  These helpers do the same so the
identical known-answer tests can run (a) on the reference's CPU build, pinning oracle/_ref,
and (b) on the same applications linked against libb200ks on the GPU box.
"""
import re


def headtail(lines, patterns, select=None):
    """headtail.pl: copy from a line matching patterns[0] through one matching patterns[1],
    then continue with the next pair; optional SELECT regex filter."""
    pats = list(patterns)
    out, start = [], False
    pa, pb = pats.pop(0), pats.pop(0)
    sel = re.compile(select) if select else None
    for ln in lines:
        if re.search(pa, ln):
            start = True
        if start:
            if sel is None or sel.search(ln):
                out.append(ln)
            if re.search(pb, ln):
                if not pats:
                    return out
                start = False
                pa, pb = pats.pop(0), pats.pop(0)
    return out


def filter_test_lines(lines):
    """Make_test_template:89: drop warnings, timing lines and comments from the test output."""
    return [ln for ln in lines if "warning" not in ln.lower() and "time =" not in ln and not ln.startswith("#")]


_NUM = re.compile(r"^[+-]?\d+(\.\d*)?([eEdDg][+-]?\d+)?$")


def _is_number(tok):
    return bool(_NUM.match(re.sub(r"[,\)]$", "", tok)))


def _val(tok):
    m = re.match(r"^[+-]?(\d+(\.\d*)?|\.\d+)([eE][+-]?\d+)?", tok.replace("D", "e").replace("d", "e"))
    return float(m.group(0)) if m else 0.0


def diffn3(test, sample, errtol):
    """diffn3.pl: returns a list of discrepancy strings (empty = OK).  Line counts must match."""
    if not (len(test) == len(sample) == len(errtol)):
        return ["line counts differ: test %d sample %d errtol %d" % (len(test), len(sample), len(errtol))]
    bad = []
    for n, (l1, l2, le) in enumerate(zip(test, sample, errtol)):
        f1, f2, errs = l1.split(), l2.split(), le.split()
        if errs and errs[0] == "XXXX":
            continue
        a, b = (f2, f1) if len(f2) > len(f1) else (f1, f2)
        for i, tok in enumerate(a):
            tol = errs[i] if i < len(errs) else "0"
            other = b[i] if i < len(b) else ""
            if tol == "XXX":
                continue
            if (not _is_number(other)) or other == "nan" or tok == "nan":
                if tok != other:
                    bad.append("line %d field %d: %r != %r" % (n + 1, i + 1, tok, other))
                continue
            diff = abs(_val(tok) - _val(other))
            if diff > _val(tol):
                bad.append("line %d field %d: diff %.3g > tol %s (%s vs %s)" % (n + 1, i + 1, diff, tol, tok, other))
    return bad
