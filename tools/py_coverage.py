#!/usr/bin/env python
"""Which Python lines of machin_b200/ does the CPU suite execute?  (No coverage package in this image.)

    python tools/py_coverage.py [pytest args...]      # default: tests -m "not gpu" -q -p no:cacheprovider

Runs pytest in-process under sys.settrace restricted to the package's files and lists, per file, the executable lines
(from the compiled code objects) that never ran.  Lines executed only inside spawned worker processes (gloo tests) show
up as missed.  A development aid for the GPU-only branches: anything listed here has never run anywhere in this container."""
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "machin_b200") + os.sep
seen = {}


def tracer(frame, event, arg):
    fn = frame.f_code.co_filename
    if not fn.startswith(PKG):
        return None
    lines = seen.setdefault(fn, set())

    def local(frame, event, arg):
        if event == "line":
            lines.add(frame.f_lineno)
        return local
    lines.add(frame.f_lineno)
    return local


def executable_lines(path):
    out = set()
    stack = [compile(open(path).read(), path, "exec")]
    while stack:
        code = stack.pop()
        out.update(ln for _, _, ln in code.co_lines() if ln is not None)
        stack.extend(c for c in code.co_consts if hasattr(c, "co_lines"))
    return out


def main():
    sys.path.insert(0, ROOT)
    import pytest
    args = sys.argv[1:] or [os.path.join(ROOT, "tests"), "-m", "not gpu", "-q", "-p", "no:cacheprovider", "-x"]
    threading.settrace(tracer)
    sys.settrace(tracer)
    rc = pytest.main(args)
    sys.settrace(None)
    total = missed_total = 0
    report = []
    for dirpath, _, files in os.walk(PKG):
        for f in sorted(files):
            if not f.endswith(".py"):
                continue
            path = os.path.join(dirpath, f)
            lines = executable_lines(path)
            src = open(path).read().splitlines()
            missed = sorted(ln for ln in lines - seen.get(path, set())
                            if src[ln - 1].strip() and not src[ln - 1].strip().startswith(('"""', "#")))
            total += len(lines)
            missed_total += len(missed)
            report.append((len(missed), os.path.relpath(path, ROOT), len(lines), missed))
    for n, path, all_lines, missed in sorted(report, reverse=True):
        if n:
            print(f"{path}: {n} of {all_lines} never ran: {missed}")
    print(f"TOTAL {total - missed_total} / {total} executable lines ran ({100.0 * (total - missed_total) / max(total, 1):.1f} %); pytest rc {rc}")


if __name__ == "__main__":
    main()
