"""TEST INFRASTRUCTURE ONLY -- a stand-in for the TensorFlow-1.x API subset the reference scripts use.

TensorFlow is not installable here (no network; the reference pins no version), so the reference's
own ``VPINN`` classes (P1D:30-224, P2D:27-257, ADI:58-341) cannot run as shipped.  This module lets
those classes execute UNMODIFIED by restating the semantics of every TF-1 symbol they call
(P1D:134-137,146-147,83-91,95,98,103-104; P2D:165-168,177-184,94-115,119-132; ADI:226-229,238-244,
162-174,181-193) on top of torch float64 CPU tensors:

* deferred graph: every ``tf.*`` op returns a :class:`Node`; nothing is computed until
  ``Session.run`` walks the graph (memoised per ``run`` call, like one TF executor step);
* ``tf.gradients(y, x)`` == gradient of ``sum(y)`` w.r.t. ``x`` (TF semantics), differentiable again;
* ``tf.train.AdamOptimizer`` == TF-1 Adam: ``lr_t = lr*sqrt(1-b2^t)/(1-b1^t)``,
  ``var -= lr_t*m/(sqrt(v)+eps)`` with ``eps=1e-8`` OUTSIDE the bias correction;
* ``tf.truncated_normal`` re-draws samples beyond two standard deviations.  The RNG stream is torch's,
  NOT TensorFlow's, so weight initialisations are not bit-reproducible against a real TF run; the
  parity fixtures therefore always inject explicit weights.

Only ``tests/`` and the fixture generator ``tests/golden/make_golden.py`` import this.  Nothing under
``hp-vpinns_b200/`` does.
"""
import sys

import numpy as np
import torch

sys.setrecursionlimit(max(sys.getrecursionlimit(), 100000))

float64 = torch.float64
float32 = torch.float32

_VARIABLES = []          # every tf.Variable ever created in this process (the "global collection")


def reset_default_graph():
    _VARIABLES.clear()


def _as_tensor(v, dtype=torch.float64):
    if isinstance(v, torch.Tensor):
        return v
    return torch.as_tensor(np.asarray(v), dtype=dtype)


class _Ctx:
    def __init__(self, feed):
        self.memo = {}
        self.feed = feed


def _ev(x, ctx):
    """Evaluate a Node / array / scalar / nested list inside one Session.run."""
    if isinstance(x, Node):
        key = id(x)
        if key not in ctx.memo:
            ctx.memo[key] = x._compute(ctx)
        return ctx.memo[key]
    if isinstance(x, (list, tuple)):
        return torch.stack([_as_tensor(_ev(e, ctx)) for e in x])
    return _as_tensor(x)


class Node:
    __array_ufunc__ = None          # make ``ndarray * Node`` defer to Node.__rmul__
    __array_priority__ = 1000

    def __init__(self, fn, inputs=(), name="op"):
        self._fn = fn
        self._inputs = tuple(inputs)
        self.name = name

    def _compute(self, ctx):
        return self._fn(*[_ev(i, ctx) for i in self._inputs])

    # arithmetic -----------------------------------------------------------
    def __add__(self, o): return Node(lambda a, b: a + b, (self, o), "add")
    def __radd__(self, o): return Node(lambda a, b: b + a, (self, o), "add")
    def __sub__(self, o): return Node(lambda a, b: a - b, (self, o), "sub")
    def __rsub__(self, o): return Node(lambda a, b: b - a, (self, o), "sub")
    def __mul__(self, o): return Node(lambda a, b: a * b, (self, o), "mul")
    def __rmul__(self, o): return Node(lambda a, b: b * a, (self, o), "mul")
    def __truediv__(self, o): return Node(lambda a, b: a / b, (self, o), "div")
    def __rtruediv__(self, o): return Node(lambda a, b: b / a, (self, o), "div")
    def __neg__(self): return Node(lambda a: -a, (self,), "neg")
    def __pow__(self, p): return Node(lambda a: a ** p, (self,), "pow")

    def __getitem__(self, idx): return Node(lambda a: a[idx], (self,), "getitem")


class _Leaf(Node):
    """A node whose per-run value is a fresh autograd leaf (constants and placeholders)."""

    def __init__(self, name):
        super().__init__(None, (), name)


class _Constant(_Leaf):
    def __init__(self, value):
        super().__init__("Const")
        self._value = value

    def _compute(self, ctx):
        return self._value.clone().requires_grad_(True)

    @property
    def shape(self):
        return tuple(self._value.shape)


class _Placeholder(_Leaf):
    def __init__(self, dtype, shape):
        super().__init__("Placeholder")
        self.dtype = dtype
        self.shape = shape

    def _compute(self, ctx):
        for k, v in ctx.feed.items():
            if k is self:
                return _as_tensor(v, self.dtype).clone().requires_grad_(True)
        raise ValueError("placeholder was not fed")


class Variable(Node):
    def __init__(self, initial_value, dtype=float64, name=None):
        super().__init__(None, (), "Variable")
        if isinstance(initial_value, Node):
            initial_value = _ev(initial_value, _Ctx({}))
        self.value = _as_tensor(initial_value, dtype).detach().clone().to(dtype).requires_grad_(True)
        _VARIABLES.append(self)

    def _compute(self, ctx):
        return self.value

    def load(self, value, session=None):
        with torch.no_grad():
            self.value.copy_(_as_tensor(value, self.value.dtype).reshape(self.value.shape))

    def assign(self, value):
        return Node(lambda: self.load(value), (), "assign")


def placeholder(dtype, shape=None, name=None):
    return _Placeholder(dtype, shape)


def constant(value, dtype=float64, shape=None, name=None):
    t = _as_tensor(value, dtype).to(dtype)
    if shape is not None:
        shp = tuple(int(s) for s in shape)
        t = t.expand(shp).clone() if t.numel() == 1 else t.reshape(shp)
    return _Constant(t)


def zeros(shape, dtype=float64):
    return _Constant(torch.zeros(tuple(shape), dtype=dtype))


def ones(shape, dtype=float64):
    return _Constant(torch.ones(tuple(shape), dtype=dtype))


def set_random_seed(seed):
    torch.manual_seed(int(seed))


def truncated_normal(shape, mean=0.0, stddev=1.0, dtype=float64, seed=None):
    def draw():
        out = torch.randn(tuple(shape), dtype=dtype)
        bad = out.abs() > 2.0
        while bool(bad.any()):
            out[bad] = torch.randn(int(bad.sum()), dtype=dtype)
            bad = out.abs() > 2.0
        return mean + float(stddev) * out
    return Node(draw, (), "truncated_normal")


def concat(values, axis):
    return Node(lambda *v: torch.cat(list(v), dim=axis), tuple(values), "concat")


def matmul(a, b): return Node(lambda x, y: x @ y, (a, b), "matmul")
def add(a, b): return Node(lambda x, y: x + y, (a, b), "add")
def tanh(a): return Node(torch.tanh, (a,), "tanh")
def sin(a): return Node(torch.sin, (a,), "sin")
def square(a): return Node(lambda x: x * x, (a,), "square")
def reduce_sum(a): return Node(torch.sum, (a,), "reduce_sum")
def reduce_mean(a): return Node(torch.mean, (a,), "reduce_mean")


def reshape(a, shape):
    return Node(lambda x: x.reshape(tuple(int(s) for s in shape)), (a,), "reshape")


def stack(values, axis=0):
    return Node(lambda *v: torch.stack([_as_tensor(e) for e in v], dim=axis), tuple(values), "stack")


def convert_to_tensor(value, dtype=None):
    if isinstance(value, Node):
        return value
    return Node(lambda: None, (), "convert")._with(lambda ctx: _ev(value, ctx))


def _with(self, compute):
    self._compute = compute
    return self


Node._with = _with


def gradients(ys, xs):
    """TF semantics: d(sum(ys))/d(xs); ``xs`` may be one tensor or a list; returns a list."""
    single = not isinstance(xs, (list, tuple))
    xs_l = [xs] if single else list(xs)

    def make(i):
        def compute(ctx):
            key = ("grads", id(ys), tuple(id(x) for x in xs_l))
            if key not in ctx.memo:
                xv = [_ev(x, ctx) for x in xs_l]
                yv = _ev(ys, ctx)
                ctx.memo[key] = torch.autograd.grad(yv, xv, grad_outputs=torch.ones_like(yv),
                                                    create_graph=True, allow_unused=True)
            return ctx.memo[key][i]
        return Node(None, (), "gradients")._with(compute)
    return [make(i) for i in range(len(xs_l))]


def global_variables_initializer():
    return Node(lambda: None, (), "init")


def trainable_variables():
    return list(_VARIABLES)


class ConfigProto:
    def __init__(self, **kw):
        self.kw = kw


class Session:
    def __init__(self, config=None):
        self.config = config

    def run(self, fetches, feed_dict=None):
        ctx = _Ctx(feed_dict or {})
        if isinstance(fetches, (list, tuple)):
            return [self._out(_ev(f, ctx)) for f in fetches]
        return self._out(_ev(fetches, ctx))

    @staticmethod
    def _out(v):
        if v is None:
            return None
        if isinstance(v, torch.Tensor):
            a = v.detach().numpy().copy()
            return a[()] if a.ndim == 0 else a
        return v

    def close(self):
        pass


class _AdamOptimizer:
    """TF-1 ``tf.train.AdamOptimizer`` (P1D:102-104, P2D:131-132, ADI:191-193)."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps = float(learning_rate), beta1, beta2, epsilon
        self.t = 0
        self.slots = {}

    def minimize(self, loss, var_list=None):
        variables = list(var_list) if var_list is not None else list(_VARIABLES)

        def step(ctx):
            lv = _ev(loss, ctx)
            grads = torch.autograd.grad(lv, [v.value for v in variables], allow_unused=True)
            self.t += 1
            lr_t = self.lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
            with torch.no_grad():
                for v, g in zip(variables, grads):
                    if g is None:       # e.g. the dead scalar `a` (P1D:117): skipped, as in TF
                        continue
                    m, s = self.slots.setdefault(id(v), (torch.zeros_like(v.value), torch.zeros_like(v.value)))
                    m.mul_(self.b1).add_(g, alpha=1 - self.b1)
                    s.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                    v.value.sub_(lr_t * m / (s.sqrt() + self.eps))
            return None
        return Node(None, (), "adam_step")._with(step)


class train:
    AdamOptimizer = _AdamOptimizer
