"""Generates the case-dictionary fixtures under tests/golden/ -- run HERE (the build container), never on the GPU box.

  plateHole.blockMeshDict.json : the parsed content of run/stressFoam/plateHole/plateHole/constant/polyMesh/blockMeshDict
  flange.polymesh.npz          : run/thermalStressFoam/flange/flange.ans through ansys.ans_to_polymesh (the case ships no polyMesh
                                 but its boundary file; 5712 cells, 15316 internal faces, 3268 boundary faces as that file says)
  <case>.case.json             : the parsed dictionaries unsTotalLagrangianStress reads (0/D, 0/pointD,
                                 constant/{rheologyProperties,stressProperties}, system/{fvSolution,fvSchemes,controlDict})
                                 of the shipped solid cases.  (Case INPUT of the reference's tutorials, not source.)

usage: python tests/golden/make_case_fixtures.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from fluidstructureinteraction_b200 import ansys, blockmesh, foamdict  # noqa: E402

RUN = "/root/reference/run"
CASES = {
    "plateHole": "stressFoam/plateHole/plateHole",
    "beam_solid": "fsiFoam/beamInCrossFlow/solid",
    "tube_solid": "fsiFoam/3dTube/solid",
    "flange": "thermalStressFoam/flange",
    "beam_fluid": "fsiFoam/beamInCrossFlow/fluid",
    "tube_fluid": "fsiFoam/3dTube/fluid",
}
FILES = {"D": "0/D", "pointD": "0/pointD", "rheologyProperties": "constant/rheologyProperties",
         "stressProperties": "constant/stressProperties", "fvSolution": "system/fvSolution", "fvSchemes": "system/fvSchemes",
         "controlDict": "system/controlDict", "p": "0/p", "T": "0/T", "thermalProperties": "constant/thermalProperties",
         "flowProperties": "constant/flowProperties"}


def main():
    d = blockmesh.read_block_mesh_dict(os.path.join(RUN, CASES["plateHole"], "constant/polyMesh/blockMeshDict"))
    with open(os.path.join(HERE, "plateHole.blockMeshDict.json"), "w") as f:
        json.dump(d, f, indent=1)
    ansys.ans_to_polymesh(os.path.join(RUN, CASES["flange"], "flange.ans")).save_npz(os.path.join(HERE, "flange.polymesh.npz"))
    for name, rel in sorted(CASES.items()):
        out = {}
        for key, path in FILES.items():
            p = os.path.join(RUN, rel, path)
            if os.path.exists(p):
                out[key] = foamdict.read(p)
        out["source"] = "run/" + rel
        with open(os.path.join(HERE, name + ".case.json"), "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)
        print(name, sorted(out))


if __name__ == "__main__":
    main()
