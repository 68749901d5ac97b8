"""
ORACLE tooling (test infrastructure): algorithmic fp64 flops per event of the integrands,
counted by running the numpy restatement (oracle/vegas_ref.py) on an ndarray subclass that
tallies every ufunc call (SURVEY.md 8d convention: add/sub/mul/div = 1, each transcendental
= 1; comparisons, selects, negation, abs of a real, casts = 0; complex ops are expanded:
complex*complex = 6, complex+-complex = 2, complex*real = 2, |complex| = 4 (2 mul, 1 add,
1 sqrt), sqrt(complex) = 1, real(complex) = 0).

    python -m oracle.count_flops
"""
import numpy as np

from oracle import vegas_ref as R

ONE = {"add", "subtract", "multiply", "divide", "true_divide", "exp", "log", "sqrt", "sin", "cos",
       "arccos", "arccosh", "sinh", "cosh", "power", "square"}
ZERO = {"negative", "absolute", "greater", "less", "equal", "not_equal", "greater_equal",
        "less_equal", "isnan", "floor", "sign", "positive", "real", "imag", "conjugate",
        "logical_and", "logical_or", "maximum", "minimum"}


class Counted(np.ndarray):
    tally = {}
    flops = 0
    zero_flops = 0
    n_events = 0

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        args = [np.asarray(a) if isinstance(a, Counted) else a for a in inputs]
        res = getattr(ufunc, method)(*args, **kwargs)
        name = ufunc.__name__
        is_c = [np.iscomplexobj(a) for a in args]
        per_event = any(getattr(np.asarray(a), "ndim", 0) >= 1 for a in args)
        cost = 0
        if per_event:
            if name in ONE:
                if any(is_c):
                    if name == "multiply":
                        cost = 6 if all(is_c) else 2
                    elif name in ("add", "subtract"):
                        cost = 2
                    elif name == "sqrt":
                        cost = 1
                    elif name in ("divide", "true_divide"):
                        cost = 2 if not is_c[1] else 11
                    else:
                        cost = 1
                else:
                    cost = 1
            elif name == "absolute" and any(is_c):
                cost = 4
            elif name not in ZERO:
                cost = 1
                Counted.tally["?" + name] = Counted.tally.get("?" + name, 0) + 1
        # terms the reference builds from exact complex zeros (czeros spinor components) are
        # not evaluated by the CUDA integrands; count them separately
        if cost and per_event and any(isinstance(a, np.ndarray) and a.ndim >= 1 and a.size
                                      and not np.any(a) for a in args):
            Counted.zero_flops += cost * (res.size // Counted.n_events
                                          if isinstance(res, np.ndarray) and res.ndim >= 1 else 0)
            cost = 0
        # an op on an [n, k] array is k operations per event
        mult = 0
        if per_event and isinstance(res, np.ndarray) and res.ndim >= 1 and Counted.n_events:
            mult = res.size // Counted.n_events
        Counted.flops += cost * mult
        Counted.tally[name] = Counted.tally.get(name, 0) + mult
        if isinstance(res, np.ndarray):
            return res.view(Counted)
        return res


def count(fn, n_dim, n=8, seed=0):
    rng = np.random.default_rng(seed)
    x = (0.05 + 0.9 * rng.random((n, n_dim))).view(Counted)
    Counted.tally, Counted.flops, Counted.zero_flops, Counted.n_events = {}, 0, 0, n
    fn(x)
    return Counted.flops, dict(Counted.tally)


if __name__ == "__main__":
    for name, d in (("symgauss", 4), ("symgauss", 8), ("symgauss", 20), ("product", 8),
                    ("drellyan_lo", 4), ("singletop_lo", 3)):
        flops, tally = count(R.INTEGRANDS[name], d)
        trans = sum(v for k, v in tally.items() if k in ("exp", "log", "sqrt", "sin", "cos",
                                                         "arccos", "arccosh", "sinh", "cosh"))
        print(f"{name:13s} d={d:2d}: integrand flops/event = {flops:5d} (+{Counted.zero_flops} on "
              f"exact-zero terms, transcendental calls {trans}); "
              f"F_alg = 12d+5+{flops} = {12 * d + 5 + flops}")
