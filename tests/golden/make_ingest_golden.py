"""Generate tests/golden/ingest_golden.json from the UNMODIFIED reference (dev container only).

For seeded synthetic PDB texts (freesasa_b200.workloads.pdb_text) the compiled reference (oracle/_ref) reads the text
with freesasa_structure_from_pdb() and everything its accessors expose — coordinates and radii as IEEE bit patterns,
labels, classes, PDB lines, residues, chains, model, classifier name — is hashed.  tests/test_ingest.py recomputes the
digest from this repo's reader; equality means byte-for-byte the same structure.

    python tests/golden/make_ingest_golden.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from freesasa_b200 import structure as st  # noqa: E402
from freesasa_b200 import workloads as w  # noqa: E402
from oracle import bindings as ob  # noqa: E402
from tests.test_ingest import OPTION_SETS, digest, snapshot  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ingest_golden.json")
TEXTS = [
    dict(n_atoms=500, seed=101),
    dict(n_atoms=700, seed=102, chains=4, hydrogens=0.25, hetatm=5, altloc=0.1, unknown=0.1),
    dict(n_atoms=400, seed=103, models=3, hydrogens=0.1, hetatm=1, element_column=False),
    dict(n_atoms=3000, seed=104, chains=6, offset=(500.0, 500.0, 500.0)),
]


def main():
    ref = st.StructureAPI(ob.ref_lib(), ob.RefResult, ob.RefParameters)
    ref.lib.freesasa_set_verbosity(2)
    cases = []
    for spec in TEXTS:
        text = w.pdb_text(**spec).encode()
        for options in OPTION_SETS:
            for classifier in (None, "naccess") if options in (0, st.INCLUDE_HETATM | st.INCLUDE_HYDROGEN) else (None,):
                s = ref.from_pdb(text, ref.classifier(classifier) if classifier else None, options)
                cases.append({"pdb_text": spec, "text_sha256": hashlib.sha256(text).hexdigest(), "options": options,
                              "classifier": classifier, "n_atoms": s.n if s else 0,
                              "digest": digest(snapshot(s)) if s else None})
    with open(OUT, "w") as f:
        json.dump({"reference": "FreeSASA 2.1.3 (oracle/_ref)", "cases": cases}, f, indent=1)
    print(f"wrote {OUT}: {len(cases)} cases")


if __name__ == "__main__":
    main()
