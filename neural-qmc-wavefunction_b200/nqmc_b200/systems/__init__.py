from .molecule import (BOHR_TO_ANGSTROM, ELEMENT_CHARGES, Molecule, hydrogen_atom, hydrogen_chain,
                       hydrogen_molecule)

__all__ = ["Molecule", "hydrogen_molecule", "hydrogen_chain", "hydrogen_atom", "ELEMENT_CHARGES", "BOHR_TO_ANGSTROM"]
