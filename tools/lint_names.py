#!/usr/bin/env python
"""Tiny undefined-name check (no pyflakes in the image): GPU-only code paths cannot be executed in the build
container, so at least every name they load must resolve to a local, an enclosing-scope, a module-level or a
builtin binding.  Usage: python tools/lint_names.py [paths...]"""
import ast
import builtins
import os
import sys


class Scope:
    def __init__(self, parent=None, is_class=False):
        self.parent, self.names, self.is_class = parent, set(), is_class

    def lookup(self, name):
        s = self
        first = True
        while s is not None:
            if (first or not s.is_class) and name in s.names:
                return True
            s, first = s.parent, False
        return False


def bound_names(node):
    """Names bound directly in this scope's body (not in nested function/class scopes)."""
    out = set()

    def visit(n, top=True):
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            out.add(n.name)
            if not top:
                return
        if isinstance(n, ast.Lambda) and not top:
            return
        if isinstance(n, (ast.ListComp, ast.SetComp, ast.DictComp, ast.GeneratorExp)) and not top:
            return
        if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            out.add(n.id)
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            for a in n.names:
                out.add((a.asname or a.name).split(".")[0])
        elif isinstance(n, ast.ExceptHandler) and n.name:
            out.add(n.name)
        elif isinstance(n, (ast.Global, ast.Nonlocal)):
            out.update(n.names)
        elif isinstance(n, ast.arg):
            out.add(n.arg)
        elif isinstance(n, ast.MatchAs) and n.name:
            out.add(n.name)
        for c in ast.iter_child_nodes(n):
            if isinstance(c, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
                out.add(c.name)
                for d in c.decorator_list:
                    visit(d, False)
                continue
            if isinstance(c, (ast.Lambda, ast.ListComp, ast.SetComp, ast.DictComp, ast.GeneratorExp)):
                # walrus targets inside comprehensions bind in the enclosing scope; ignore that corner
                continue
            visit(c, False)

    visit(node)
    return out


def check(tree, filename):
    problems = []
    builtin = set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__builtins__", "__spec__", "__package__"}

    def walk(node, scope):
        for child in ast.iter_child_nodes(node):
            if isinstance(child, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
                if not isinstance(child, ast.Lambda):
                    for d in child.decorator_list:
                        walk_expr(d, scope)
                    if child.returns:
                        pass
                for default in child.args.defaults + [d for d in child.args.kw_defaults if d is not None]:
                    walk_expr(default, scope)
                inner = Scope(scope)
                inner.names = bound_names(child)
                walk(child, inner) if not isinstance(child, ast.Lambda) else walk_expr(child.body, inner)
            elif isinstance(child, ast.ClassDef):
                for b in child.bases:
                    walk_expr(b, scope)
                inner = Scope(scope, is_class=True)
                inner.names = bound_names(child)
                walk(child, inner)
            elif isinstance(child, (ast.ListComp, ast.SetComp, ast.DictComp, ast.GeneratorExp)):
                inner = Scope(scope)
                for gen in child.generators:
                    for n in ast.walk(gen.target):
                        if isinstance(n, ast.Name):
                            inner.names.add(n.id)
                walk(child, inner)
            elif isinstance(child, ast.Name):
                if isinstance(child.ctx, ast.Load) and not scope.lookup(child.id) and child.id not in builtin:
                    problems.append((filename, child.lineno, child.id))
            elif isinstance(child, ast.arguments):
                for a in child.args + child.kwonlyargs + child.posonlyargs:
                    if a.annotation is not None:
                        pass     # annotations may be strings / forward references
            else:
                walk(child, scope)

    def walk_expr(expr, scope):
        holder = ast.Expr(value=expr)
        walk(holder, scope)

    top = Scope()
    top.names = bound_names(tree)
    walk(tree, top)
    return problems


def main():
    paths = sys.argv[1:] or ["machin_b200", "oracle", "tests", "tools", "bench.py", "bench_workloads.py", "__graft_entry__.py"]
    files = []
    for p in paths:
        if os.path.isdir(p):
            for root, dirs, names in os.walk(p):
                dirs[:] = [d for d in dirs if d != "_ref"]       # oracle/_ref: the vendored, unmodified reference (not ours)
                files += [os.path.join(root, n) for n in names if n.endswith(".py")]
        else:
            files.append(p)
    bad = []
    for f in sorted(files):
        bad += check(ast.parse(open(f).read(), f), f)
    for f, line, name in bad:
        print(f"{f}:{line}: undefined name {name!r}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
