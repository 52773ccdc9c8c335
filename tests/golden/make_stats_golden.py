"""Generates tests/golden/stats_golden.json by EXECUTING the reference's [SPEC] statistics starter code
(/root/reference/.github/phase5-issue-body.md:112-151 -- plain NumPy, runs as written) on seeded AR(1) series.
Build container only:  python tests/golden/make_stats_golden.py"""
import json
import os
from typing import Tuple  # noqa: F401  (used by the extracted code)

import numpy as np

MD, LO, HI = "/root/reference/.github/phase5-issue-body.md", 112, 151
lines = open(MD, encoding="utf-8").read().split("\n")
assert lines[LO - 2].startswith("```python") and lines[HI].startswith("```")
ns = {"np": np, "Tuple": Tuple}
exec(compile("\n".join(lines[LO - 1:HI]), f"{MD}:{LO}-{HI}", "exec"), ns)

out = {"source": f"{MD}:{LO}-{HI}", "cases": []}
for seed, n, rho in [(0, 4000, 0.0), (1, 4000, 0.7), (2, 3001, 0.95), (3, 257, 0.5)]:
    rng = np.random.default_rng(seed)
    x = np.empty(n)
    x[0] = rng.standard_normal()
    for i in range(1, n):
        x[i] = rho * x[i - 1] + np.sqrt(1 - rho * rho) * rng.standard_normal()
    x = x - 1.17
    c = {"seed": seed, "n": n, "rho": rho}
    for nb in (20, 7):
        m, s = ns["blocking_analysis"](x, nb)
        c[f"blocking_{nb}"] = [float(m), float(s)]
    c["acf_first8"] = [float(v) for v in ns["compute_autocorrelation"](x, 8)]
    c["ess"] = float(ns["effective_sample_size"](x))
    out["cases"].append(c)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "stats_golden.json"), "w"), indent=1)
print("wrote", len(out["cases"]), "cases")
