#!/usr/bin/env python
"""Generates tests/golden/intree_models.json by RUNNING the reference's own in-tree Python models
(/root/reference/microscopes/dbg/models/bbnc.py and dm.py) -- the only implementations of a family on the hot
path that the reference tree itself holds in an importable language.

Run in the BUILD container only (it imports from /root/reference); the JSON is committed.

The two files are Python-2 era and import three helper modules of the un-vendored `distributions` package; these
are shimmed with what their names say (nothing of the models themselves is touched):
  distributions.dbg.special   log = math.log, gammaln = scipy.special.gammaln
  distributions.dbg.random    sample_beta / sample_bernoulli (numpy; only used at Group.init, whose draw is
                              overwritten with a fixed p below)
  distributions.mixins        empty base classes
  numpy.float / numpy.int     aliases removed in numpy 1.24

What is recorded: bbnc Group.add_value / remove_value / score_value / score_data (bbnc.py:63-92) and dm
Group.add_value / remove_value (dm.py:39-53: counts and the multinomial-coefficient `ratio`).  dm.py's score_value
raises "need to fix" (:59-70) and its score_data reads an attribute that does not exist (`self._alphas`, :75), so
for those two the C++ source (src/models/dm.cpp:38-95) stays the only statement; they are pinned against scipy in
score_value.json instead.
"""
import importlib.util
import json
import math
import os
import sys
import types

import numpy as np
import scipy
import scipy.special
import scipy.stats  # noqa: F401  (bbnc.py uses sp.stats through `import scipy as sp`)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/microscopes/dbg/models"


def load(name):
    for alias, typ in (("float", float), ("int", int)):
        if not hasattr(np, alias):
            setattr(np, alias, typ)
    rng = np.random.default_rng(0)
    shims = {
        "distributions": types.ModuleType("distributions"),
        "distributions.dbg": types.ModuleType("distributions.dbg"),
        "distributions.dbg.special": types.ModuleType("distributions.dbg.special"),
        "distributions.dbg.random": types.ModuleType("distributions.dbg.random"),
        "distributions.mixins": types.ModuleType("distributions.mixins"),
    }
    shims["distributions.dbg.special"].log = math.log
    shims["distributions.dbg.special"].gammaln = lambda x: float(scipy.special.gammaln(x))
    shims["distributions.dbg.random"].sample_beta = lambda a, b: float(rng.beta(a, b))
    shims["distributions.dbg.random"].sample_bernoulli = lambda p: bool(rng.random() < p)
    for cls in ("SharedMixin", "GroupIoMixin", "SharedIoMixin"):
        setattr(shims["distributions.mixins"], cls, type(cls, (object,), {}))
    sys.modules.update(shims)
    spec = importlib.util.spec_from_file_location("ref_dbg_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    bbnc, dm = load("bbnc"), load("dm")
    rng = np.random.default_rng(20141209)
    out = {"generator": "scripts/make_golden_intree.py", "bbnc": [], "dm": []}

    # ---- bbnc: the module's own EXAMPLES entry first, then random sequences ----
    runs = [(e["shared"]["alpha"], e["shared"]["beta"], 0.3, [bool(v) for v in e["values"]]) for e in bbnc.EXAMPLES]
    for _ in range(6):
        runs.append((float(rng.uniform(0.3, 4)), float(rng.uniform(0.3, 4)), float(rng.uniform(0.02, 0.98)),
                     [bool(v) for v in rng.integers(0, 2, size=int(rng.integers(0, 40)))]))
    for alpha, beta, p, values in runs:
        shared = bbnc.Shared()
        shared.load({"alpha": alpha, "beta": beta})
        g = bbnc.Group()
        g.init(shared)
        g.p = p
        for v in values:
            g.add_value(shared, v)
        removed = values[: len(values) // 3]
        rec = dict(alpha=alpha, beta=beta, p=p, values=[int(v) for v in values], after_add=[g.heads, g.tails],
                   score_true=g.score_value(shared, True), score_false=g.score_value(shared, False),
                   score_data=float(g.score_data(shared)))
        for v in removed:
            g.remove_value(shared, v)
        rec.update(removed=[int(v) for v in removed], after_remove=[g.heads, g.tails], score_data_after_remove=float(g.score_data(shared)))
        out["bbnc"].append(rec)

    # ---- dm: add / remove bookkeeping ----
    for C, nrows, tot in [(3, 5, 6), (8, 20, 30), (16, 40, 200), (1, 4, 9), (5, 0, 0)]:
        shared = dm.Shared()
        shared.load({"alphas": [1.0] * C})
        shared.alphas = shared._alphas   # dm.py:20 reads `self.alphas`, which dm.py never sets (load() fills `_alphas`)
        g = dm.Group()
        g.init(shared)
        rows = [rng.multinomial(int(rng.integers(0, tot + 1)), rng.dirichlet(np.ones(C))).tolist() for _ in range(nrows)]
        for x in rows:
            g.add_value(shared, list(enumerate(x)))          # dm.py:41 iterates (index, count) pairs
        rec = dict(dim=C, rows=rows, counts_after_add=g._counts.tolist(), ratio_after_add=float(g._ratio))
        removed = rows[: nrows // 2]
        for x in removed:
            g.remove_value(shared, list(enumerate(x)))
        rec.update(removed=len(removed), counts_after_remove=g._counts.tolist(), ratio_after_remove=float(g._ratio))
        out["dm"].append(rec)

    path = os.path.join(ROOT, "tests", "golden", "intree_models.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote %d bbnc runs and %d dm runs to %s" % (len(out["bbnc"]), len(out["dm"]), path))


if __name__ == "__main__":
    main()
