#!/usr/bin/env python
"""Poor man's pyflakes (none is installed here): names that are read somewhere in a function but bound nowhere in
its scope chain, the module or builtins.  Usage: python tools/check_names.py file.py ..."""
import ast
import builtins
import sys


class Scope:
    def __init__(self, node, parent):
        self.node, self.parent, self.bound, self.loads = node, parent, set(), []


def bind_target(scope, t):
    for n in ast.walk(t):
        if isinstance(n, ast.Name):
            scope.bound.add(n.id)


def visit(node, scope, scopes):
    if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda, ast.ClassDef)):
        if not isinstance(node, ast.Lambda):
            scope.bound.add(node.name)
            for d in node.decorator_list:
                visit(d, scope, scopes)
        inner = Scope(node, scope)
        scopes.append(inner)
        if not isinstance(node, ast.ClassDef):
            a = node.args
            for x in a.posonlyargs + a.args + a.kwonlyargs + ([a.vararg] if a.vararg else []) + ([a.kwarg] if a.kwarg else []):
                inner.bound.add(x.arg)
            for d in a.defaults + [d for d in a.kw_defaults if d is not None]:
                visit(d, scope, scopes)
        body = node.body if isinstance(node.body, list) else [node.body]
        for b in body:
            visit(b, inner, scopes)
        return
    if isinstance(node, (ast.ListComp, ast.SetComp, ast.DictComp, ast.GeneratorExp)):
        inner = Scope(node, scope)
        scopes.append(inner)
        for g in node.generators:
            bind_target(inner, g.target)
        for ch in ast.iter_child_nodes(node):
            visit(ch, inner, scopes)
        return
    if isinstance(node, ast.Name):
        if isinstance(node.ctx, ast.Load):
            scope.loads.append((node.id, node.lineno))
        else:
            scope.bound.add(node.id)
    elif isinstance(node, (ast.Import, ast.ImportFrom)):
        for al in node.names:
            scope.bound.add((al.asname or al.name).split(".")[0])
    elif isinstance(node, ast.ExceptHandler) and node.name:
        scope.bound.add(node.name)
    elif isinstance(node, (ast.Global, ast.Nonlocal)):
        scope.bound.update(node.names)
    elif isinstance(node, ast.NamedExpr):
        bind_target(scope, node.target)
    for ch in ast.iter_child_nodes(node):
        visit(ch, scope, scopes)


def check(path):
    tree = ast.parse(open(path).read(), path)
    top = Scope(tree, None)
    scopes = [top]
    for b in tree.body:
        visit(b, top, scopes)
    bad = 0
    for s in scopes:
        for name, line in s.loads:
            p = s
            while p is not None and name not in p.bound:
                p = p.parent
                while p is not None and isinstance(p.node, ast.ClassDef):      # class scopes are not enclosing scopes
                    p = p.parent
            if p is None and not hasattr(builtins, name) and name not in ("__file__", "__name__", "__doc__"):
                print("%s:%d: undefined name %r" % (path, line, name))
                bad += 1
    return bad


if __name__ == "__main__":
    sys.exit(1 if sum(check(p) for p in sys.argv[1:]) else 0)
