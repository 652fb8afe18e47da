"""Generate the packaged input fixtures from the reference's ``utilities/`` folder.

Run in the build container (where /root/reference exists):
    python tools/make_data.py [/root/reference/utilities]
Writes
    gmpnp_b200/data/reference_inputs.json   parsed parameters*.yaml + bulk_soln_*KHCO3.yaml
    gmpnp_b200/data/meshes/<stem>.npz        vertices/cells of every dolfin XML mesh
These are data (physical constants, mesh coordinates), needed on machines where
the reference checkout is absent (GPU box).  No reference code is copied.
"""
import glob
import json
import os
import sys

import yaml

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gmpnp_b200 import meshio  # noqa: E402


def main(util="/root/reference/utilities"):
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gmpnp_b200", "data")
    os.makedirs(os.path.join(out, "meshes"), exist_ok=True)
    inputs = {}
    for p in sorted(glob.glob(os.path.join(util, "*.yaml"))):
        with open(p) as f:
            inputs[os.path.basename(p)[:-5]] = yaml.safe_load(f)
    with open(os.path.join(out, "reference_inputs.json"), "w") as f:
        json.dump(inputs, f, indent=1, sort_keys=True)
    seen = {}
    for p in sorted(glob.glob(os.path.join(util, "*.xml")) + glob.glob(os.path.join(util, "*.xml.gz"))):
        m = meshio.read_dolfin_xml(p)
        meshio.save_npz(m, os.path.join(out, "meshes", m.name + ".npz"))
        print(m.name, m.num_vertices, m.num_cells)
        seen[m.name] = (m.num_vertices, m.num_cells)
    with open(os.path.join(out, "meshes", "index.json"), "w") as f:
        json.dump(seen, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main(*sys.argv[1:])
