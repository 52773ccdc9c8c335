"""Generates tests/golden/molecule_golden.json by EXECUTING the reference's own
src/nqmc/systems/molecule.py (the only reference module importable without JAX; loaded by
file path because nqmc/__init__.py pulls jax).  Run in the build container only:

    python tests/golden/make_molecule_golden.py

/root/reference does not exist on the GPU box; tests read the committed JSON.
"""
import importlib.util, json, os
import numpy as np

REF = "/root/reference/src/nqmc/systems/molecule.py"
spec = importlib.util.spec_from_file_location("ref_molecule", REF)
m = importlib.util.module_from_spec(spec)
import sys
sys.modules["ref_molecule"] = m      # dataclasses needs the module registered
spec.loader.exec_module(m)

out = {"source": REF, "cases": []}


def dump(name, mol, probe=None):
    c = {"name": name, "n_atoms": mol.n_atoms, "n_electrons": mol.n_electrons, "spin": list(mol.spin),
         "positions": mol.positions.tolist(), "charges": mol.charges.tolist(),
         "v_nn": float(mol.nuclear_repulsion_energy()),
         "spin_mask": mol.get_spin_mask().astype(int).tolist(),
         "xyz": mol.to_xyz_string()}
    if probe is not None:
        c["probe_r"] = probe.tolist()
        c["r_en"] = mol.electron_nuclear_distances(probe).tolist()
        c["r_ee"] = mol.electron_electron_distances(probe).tolist()
    out["cases"].append(c)


rng = np.random.default_rng(1234)
dump("h_atom", m.hydrogen_atom(), rng.standard_normal((1, 3)))
dump("h2_1.4", m.hydrogen_molecule(1.4), np.array([[0.5, 0.0, 0.0], [1.0, 0.0, 0.0]]))
for n in (2, 4, 6, 8, 10):
    mol = m.hydrogen_chain(n, 1.8)
    dump(f"h{n}_chain_1.8", mol, rng.standard_normal((mol.n_electrons, 3)) + mol.positions[np.arange(mol.n_electrons) % n])
for R in np.linspace(1.0, 4.0, 16):
    dump(f"h4_sweep_{R:.2f}", m.hydrogen_chain(4, float(R)))
dump("lih", m.Molecule.from_atoms(["Li", "H"], np.array([[0, 0, 0], [3.015, 0, 0]])), rng.standard_normal((4, 3)))
dump("h3_plus_like", m.Molecule.from_atoms(["H", "H", "H"], np.array([[0, 0, 0], [1.6, 0, 0], [0.8, 1.4, 0]]), charge=1))
dump("be_triplet", m.Molecule.from_atoms(["Be"], np.zeros((1, 3)), spin_polarization=2))
errs = {}
try:
    m.hydrogen_chain(3)
except ValueError as e:
    errs["odd_chain"] = "ValueError"
try:
    m.Molecule.from_atoms(["H", "H"], np.zeros((3, 3)))
except ValueError as e:
    errs["length_mismatch"] = "ValueError"
try:
    m.Molecule.from_atoms(["Xx"], np.zeros((1, 3)))
except KeyError as e:
    errs["unknown_element"] = "KeyError"
try:
    m.Molecule.from_atoms(["H"], np.zeros((1, 3)), spin_polarization=-3)
except ValueError as e:
    errs["bad_spin"] = "ValueError"
out["errors"] = errs
out["constants"] = {"BOHR_TO_ANGSTROM": m.BOHR_TO_ANGSTROM, "ELEMENT_CHARGES": m.ELEMENT_CHARGES}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "molecule_golden.json")
json.dump(out, open(path, "w"), indent=1)
print("wrote", path, len(out["cases"]), "cases", errs)
