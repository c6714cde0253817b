"""Static hygiene for every Python file the driver executes: an undefined global name (e.g. a missing `import os`) in a
GPU-only test would otherwise surface for the first time on the B200 box.  No linter is installed in this image, so this
is a small scope-aware pass over the AST: every name that is loaded somewhere in a module must be bound somewhere in an
enclosing scope of that module (assignment, import, def, class, argument, for / with / except target, comprehension
variable, global / nonlocal declaration) or be a builtin."""
import ast
import builtins
import glob
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = sorted(glob.glob(os.path.join(ROOT, "tests", "*.py")) + glob.glob(os.path.join(ROOT, "tests", "golden", "*.py")) +
               glob.glob(os.path.join(ROOT, "anticipated-vins-mono_b200", "*.py")) +
               glob.glob(os.path.join(ROOT, "tools", "*.py")) + glob.glob(os.path.join(ROOT, "adapters", "**", "*.py"),
                                                                             recursive=True) +
               [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")])
BUILTINS = set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__spec__", "__path__", "__package__"}


class Scope:
    def __init__(self, parent, kind):
        self.parent, self.kind, self.bound, self.loads = parent, kind, set(), []


class Binder(ast.NodeVisitor):
    """Collects bindings per scope and the loads made in each scope."""

    def __init__(self):
        self.scope = Scope(None, "module")
        self.all_scopes = [self.scope]
        self.star_import = False

    def push(self, kind):
        s = Scope(self.scope, kind)
        self.all_scopes.append(s)
        self.scope = s
        return s

    def pop(self):
        self.scope = self.scope.parent

    def bind(self, name):
        self.scope.bound.add(name)

    def visit_Import(self, node):
        for a in node.names:
            self.bind((a.asname or a.name).split(".")[0])

    def visit_ImportFrom(self, node):
        for a in node.names:
            if a.name == "*":
                self.star_import = True
            else:
                self.bind(a.asname or a.name)

    def visit_Global(self, node):
        for n in node.names:
            self.bind(n)
            self.all_scopes[0].bound.add(n)

    visit_Nonlocal = visit_Global

    def visit_Name(self, node):
        if isinstance(node.ctx, ast.Load):
            self.scope.loads.append((node.id, node.lineno))
        else:
            self.bind(node.id)

    def _args(self, args):
        for a in args.posonlyargs + args.args + args.kwonlyargs:
            self.bind(a.arg)
        if args.vararg:
            self.bind(args.vararg.arg)
        if args.kwarg:
            self.bind(args.kwarg.arg)

    def _function(self, node):
        for d in node.decorator_list:
            self.visit(d)
        for d in node.args.defaults + [k for k in node.args.kw_defaults if k is not None]:
            self.visit(d)
        for a in node.args.posonlyargs + node.args.args + node.args.kwonlyargs:
            if a.annotation is not None:
                self.visit(a.annotation)
        if node.returns is not None:
            self.visit(node.returns)
        self.bind(node.name)
        self.push("function")
        self._args(node.args)
        for st in node.body:
            self.visit(st)
        self.pop()

    visit_FunctionDef = _function
    visit_AsyncFunctionDef = _function

    def visit_Lambda(self, node):
        for d in node.args.defaults + [k for k in node.args.kw_defaults if k is not None]:
            self.visit(d)
        self.push("function")
        self._args(node.args)
        self.visit(node.body)
        self.pop()

    def visit_ClassDef(self, node):
        for d in node.decorator_list + node.bases + [k.value for k in node.keywords]:
            self.visit(d)
        self.bind(node.name)
        self.push("class")
        for st in node.body:
            self.visit(st)
        self.pop()

    def _comp(self, node):
        self.push("function")
        for g in node.generators:
            self.visit(g.iter)
            self.visit(g.target)
            for c in g.ifs:
                self.visit(c)
        if isinstance(node, ast.DictComp):
            self.visit(node.key)
            self.visit(node.value)
        else:
            self.visit(node.elt)
        self.pop()

    visit_ListComp = visit_SetComp = visit_GeneratorExp = visit_DictComp = _comp

    def visit_ExceptHandler(self, node):
        if node.name:
            self.bind(node.name)
        self.generic_visit(node)

    def visit_MatchAs(self, node):
        if node.name:
            self.bind(node.name)
        self.generic_visit(node)

    def visit_MatchStar(self, node):
        if node.name:
            self.bind(node.name)


def undefined_names(path):
    tree = ast.parse(open(path).read(), path)
    b = Binder()
    b.visit(tree)
    if b.star_import:
        return []
    bad = []
    for s in b.all_scopes:
        for name, line in s.loads:
            t, found = s, False
            while t is not None:
                # class bodies are not enclosing scopes for nested functions, but are for their own loads
                if name in t.bound and (t is s or t.kind != "class"):
                    found = True
                    break
                t = t.parent
            if not found and name not in BUILTINS:
                bad.append(f"{os.path.relpath(path, ROOT)}:{line}: undefined name '{name}'")
    return bad


@pytest.mark.parametrize("path", FILES, ids=[os.path.relpath(f, ROOT) for f in FILES])
def test_no_undefined_names(path):
    assert undefined_names(path) == []


def test_checker_catches_a_missing_import(tmp_path):
    f = tmp_path / "x.py"
    f.write_text("import sys\n\ndef g():\n    return os.path.join(sys.argv[0], 'a')\n")
    assert any("'os'" in m for m in undefined_names(str(f)))


def test_gpu_tests_collect():
    """`pytest --collect-only -m gpu` must import every GPU test module on the CPU box."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "--collect-only", "-q", "-m", "gpu"],
                       capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
