"""Host geometry layer and the oracle's geometry vs values produced by executing the reference's own
src/nqmc/systems/molecule.py (tests/golden/molecule_golden.json, made by tests/golden/make_molecule_golden.py)."""
import json
import os

import numpy as np
import pytest

from nqmc_b200.systems import molecule as M
from oracle.layout import hydrogen_chain_system

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "molecule_golden.json")))


def _build(name):
    if name == "h_atom":
        return M.hydrogen_atom()
    if name.endswith("_chain_1.8"):
        return M.hydrogen_chain(int(name[1:name.index("_")]), 1.8)
    if name == "h2_1.4":
        return M.hydrogen_molecule(1.4)
    if name.startswith("h4_sweep_"):
        return None
    if name == "lih":
        return M.Molecule.from_atoms(["Li", "H"], np.array([[0, 0, 0], [3.015, 0, 0]]))
    if name == "h3_plus_like":
        return M.Molecule.from_atoms(["H", "H", "H"], np.array([[0, 0, 0], [1.6, 0, 0], [0.8, 1.4, 0]]), charge=1)
    if name == "be_triplet":
        return M.Molecule.from_atoms(["Be"], np.zeros((1, 3)), spin_polarization=2)
    raise KeyError(name)


@pytest.mark.parametrize("case", G["cases"], ids=[c["name"] for c in G["cases"]])
def test_molecule_matches_reference(case):
    mol = _build(case["name"])
    if mol is None:
        mol = M.hydrogen_chain(4, np.array(case["positions"])[1, 0])
    assert mol.n_atoms == case["n_atoms"] and mol.n_electrons == case["n_electrons"]
    assert list(mol.spin) == case["spin"]
    assert np.array_equal(mol.positions, np.array(case["positions"]))
    assert np.array_equal(mol.charges, np.array(case["charges"]))
    assert mol.nuclear_repulsion_energy() == pytest.approx(case["v_nn"], rel=1e-15, abs=0)
    assert np.array_equal(mol.get_spin_mask().astype(int), np.array(case["spin_mask"]))
    assert mol.to_xyz_string() == case["xyz"]
    if "probe_r" in case:
        r = np.array(case["probe_r"])
        assert np.allclose(mol.electron_nuclear_distances(r), np.array(case["r_en"]), rtol=1e-15, atol=1e-15)
        assert np.allclose(mol.electron_electron_distances(r), np.array(case["r_ee"]), rtol=1e-15, atol=1e-15)


def test_golden_constants_of_survey_appendix_f():
    by = {c["name"]: c for c in G["cases"]}
    assert by["h2_1.4"]["v_nn"] == pytest.approx(0.7142857142857143)
    assert by["h4_chain_1.8"]["v_nn"] == pytest.approx(2.4074074074074074)
    assert by["h6_chain_1.8"]["v_nn"] == pytest.approx(4.833333333333333)
    assert by["h10_chain_1.8"]["v_nn"] == pytest.approx(10.716490299823633)
    assert by["h2_1.4"]["r_ee"][0][1] == pytest.approx(0.5)
    for c in G["cases"]:
        if c["name"].startswith("h4_sweep_"):
            R = c["positions"][1][0]
            assert c["v_nn"] == pytest.approx(13.0 / (3.0 * R), rel=1e-12)


def test_errors_match_reference():
    assert G["errors"] == {"odd_chain": "ValueError", "length_mismatch": "ValueError", "unknown_element": "KeyError",
                           "bad_spin": "ValueError"}
    with pytest.raises(ValueError):
        M.hydrogen_chain(3)
    with pytest.raises(ValueError):
        M.Molecule.from_atoms(["H", "H"], np.zeros((3, 3)))
    with pytest.raises(KeyError):
        M.Molecule.from_atoms(["Xx"], np.zeros((1, 3)))
    with pytest.raises(ValueError):
        M.Molecule.from_atoms(["H"], np.zeros((1, 3)), spin_polarization=-3)
    assert M.BOHR_TO_ANGSTROM == G["constants"]["BOHR_TO_ANGSTROM"]
    assert M.ELEMENT_CHARGES == G["constants"]["ELEMENT_CHARGES"]


@pytest.mark.parametrize("n", [1, 2, 4, 6, 10])
def test_oracle_geometry_matches_reference(n):
    name = "h_atom" if n == 1 else f"h{n}_chain_1.8"
    c = {x["name"]: x for x in G["cases"]}[name]
    s = hydrogen_chain_system(n, 1.8)
    assert np.array_equal(s.R, np.array(c["positions"]))
    assert [s.n_up, s.n_dn] == c["spin"]
    assert s.v_nn == pytest.approx(c["v_nn"], rel=1e-15)
