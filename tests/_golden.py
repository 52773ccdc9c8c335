"""Loader for tests/golden/ref_hotpath_*.npz -- inputs/outputs produced by EXECUTING the reference's own
source files (tests/golden/make_ref_hotpath_golden.py, oracle/_ref_shim).  Works without /root/reference."""
import json
import os

import numpy as np

from oracle.layout import KIND_ANTISYMMETRIC, KIND_SIMPLE, OracleNet, OracleSystem

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MANIFEST = json.load(open(os.path.join(GOLD, "ref_hotpath_manifest.json")))
CASE_NAMES = sorted(MANIFEST["cases"])
SIMPLE_CASES = [n for n in CASE_NAMES if MANIFEST["cases"][n]["kind"] == "simple"]
SPEC_CASES = [n for n in CASE_NAMES if MANIFEST["cases"][n]["kind"] == "antisymmetric"]


def load(name):
    """-> (meta, arrays, OracleSystem, OracleNet)"""
    meta = MANIFEST["cases"][name]
    arr = dict(np.load(os.path.join(GOLD, f"ref_hotpath_{name}.npz")))
    mol = meta["molecule"]
    s = OracleSystem(R=np.asarray(mol["positions"], dtype=np.float64), Z=np.asarray(mol["charges"], dtype=np.float64),
                     n_up=int(mol["spin"][0]), n_dn=int(mol["spin"][1]), v_nn=float(meta["V_nn"]))
    if meta["kind"] == "simple":
        net = OracleNet(kind=KIND_SIMPLE, hidden=tuple(meta["hidden"]), activation=meta["activation"],
                        envelope_decay=float(meta.get("envelope_decay", 1.0)))
    else:
        net = OracleNet(kind=KIND_ANTISYMMETRIC, hidden=tuple(meta["hidden"]), use_cusp=bool(meta["use_cusp"]),
                        use_backflow=bool(meta["use_backflow"]))
    return meta, arr, s, net
