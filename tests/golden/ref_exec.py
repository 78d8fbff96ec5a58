"""Execute the reference's OWN Python function bodies on torch 2.x (build container only; needs /root/reference).

lib/sub_module.py, lib/model.py and tools/utils.py cannot be imported (torch.utils.ffi extensions, `past`, matplotlib,
instance-style autograd Functions -- SURVEY.md Appendix C), so the functions on the hot path are cut out of the source
with `ast` -- by name, or by line range for the level rule, which is inline code of Dev.forward -- and exec'd
UNMODIFIED inside `torch03()`, a context that restores the PyTorch-0.3 behaviours those bodies rely on:

* no 0-dim tensors: indexing / squeeze / full reductions give 1-element 1-D tensors (`window[0].data[0]`, `_idx.size(0)`);
* comparisons give uint8 "ByteTensor"s, so `(a > 0) + (b > 0) == 2` (lib/model.py:180-181) means AND;
* `x in tensor` is `(x == tensor).any()` for any x, numpy arrays included (lib/model.py:173);
* `buf[:-1] = buf[1:]` (lib/model.py:161-164) is a front-to-back element copy, i.e. a shift: the value is cloned first
  (torch 2.x refuses overlapping in-place copies);
* `Variable(x, ...)` is x, `.cuda()` is the identity (this container has no GPU).

Nothing here is imported by the tests or the product: it only feeds tests/golden/make_golden.py.
"""
import ast
import contextlib
import os
import textwrap

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"


def _keep1(t):
    return t.view(1) if isinstance(t, torch.Tensor) and t.dim() == 0 else t


@contextlib.contextmanager
def torch03():
    T = torch.Tensor
    saved = {n: getattr(T, n) for n in ("__getitem__", "squeeze", "cuda", "__gt__", "__lt__", "__ge__", "__le__", "__eq__", "__ne__", "__hash__", "__contains__", "__setitem__")}
    getitem, squeeze = T.__getitem__, T.squeeze

    def _cmp(name):
        f = saved[name]

        def op(self, other):
            r = f(self, other)
            return r.to(torch.uint8) if isinstance(r, torch.Tensor) else r
        return op

    T.__getitem__ = lambda self, idx: _keep1(getitem(self.view(1) if self.dim() == 0 else self, idx))
    T.squeeze = lambda self, *a, **k: _keep1(squeeze(self, *a, **k))
    T.cuda = lambda self, *a, **k: self
    setitem = T.__setitem__
    T.__setitem__ = lambda self, idx, val: setitem(self, idx, val.clone() if isinstance(val, torch.Tensor) else val)
    for n in ("__gt__", "__lt__", "__ge__", "__le__", "__eq__", "__ne__"):
        setattr(T, n, _cmp(n))
    T.__hash__ = lambda self: id(self)          # defining __eq__ on a class drops its hash
    T.__contains__ = lambda self, el: bool(saved["__eq__"](self, torch.as_tensor(el).to(self.dtype)).any())
    try:
        yield
    finally:
        for n, f in saved.items():
            setattr(T, n, f)


def Variable(x, requires_grad=False, volatile=False):
    return x


def _source(rel):
    return open(os.path.join(REF, rel)).read()


def extract_function(rel, name, cls=None):
    """Source text of function `name` (inside class `cls` if given) of the reference file `rel`, dedented, decorators dropped."""
    src = _source(rel)
    tree = ast.parse(src)
    scope = tree.body
    if cls is not None:
        scope = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    node = next(n for n in scope if isinstance(n, ast.FunctionDef) and n.name == name)
    lines = src.split("\n")[node.lineno - 1:node.end_lineno]
    return textwrap.dedent("\n".join(lines)), (node.lineno, node.end_lineno)


def extract_lines(rel, first, last):
    """Lines first..last (1-based, inclusive) of the reference file, dedented -- for inline code of a long method."""
    return textwrap.dedent("\n".join(_source(rel).split("\n")[first - 1:last]))


def namespace(**extra):
    ns = dict(torch=torch, np=np, F=F, Variable=Variable, EPS=1e-20)
    ns.update(extra)
    return ns


def load_functions(specs, **extra):
    """specs: [(rel, name, cls)] -> namespace with those functions defined from the reference's source text."""
    ns = namespace(**extra)
    cited = {}
    for rel, name, cls in specs:
        text, span = extract_function(rel, name, cls)
        exec(compile(text, os.path.join(REF, rel), "exec"), ns)
        cited[name] = "%s:%d-%d" % (rel, span[0], span[1])
    ns["__cited__"] = cited
    return ns
