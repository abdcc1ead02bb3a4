"""Extract real pocket geometry for the parity fixtures (dev container only; SURVEY.md §8d "geometry option A").

    python scripts/make_real_geometry.py

Plain-text parse of the reference's inference examples (FABind/inference_examples/pdb_files/*.pdb and
gt_mol_files/*/*.sdf): CA coordinates of every residue, heavy-atom coordinates and heavy-atom bonds of the
ligand.  Writes tests/golden/real_geometry.npz (a few KB), which travels to the GPU box; the reference tree does not.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = "/root/reference/FABind/inference_examples"
IDS = ["6efk", "6g3c", "6n93", "6npi"]


def parse_ca(path):
    seen, out = set(), []
    for line in open(path):
        if line.startswith("ENDMDL"):
            break
        if line.startswith("ATOM") and line[12:16].strip() == "CA" and line[16] in " A":
            key = (line[21], line[22:27])
            if key in seen:
                continue
            seen.add(key)
            out.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
    return np.asarray(out, dtype=np.float32)


def parse_sdf(path):
    lines = open(path).read().splitlines()
    na, nb = int(lines[3][0:3]), int(lines[3][3:6])
    xyz, elem = [], []
    for l in lines[4:4 + na]:
        xyz.append([float(l[0:10]), float(l[10:20]), float(l[20:30])])
        elem.append(l[31:34].strip())
    heavy = [i for i, e in enumerate(elem) if e != "H"]
    remap = {a: i for i, a in enumerate(heavy)}
    bonds = []
    for l in lines[4 + na:4 + na + nb]:
        a, b = int(l[0:3]) - 1, int(l[3:6]) - 1
        if a in remap and b in remap:
            bonds += [[remap[a], remap[b]], [remap[b], remap[a]]]
    return np.asarray(xyz, dtype=np.float32)[heavy], np.asarray(sorted(bonds), dtype=np.int64).T


def main():
    out = {}
    for pid in IDS:
        ca = parse_ca(os.path.join(EX, "pdb_files", pid + ".pdb"))
        lig, bonds = parse_sdf(os.path.join(EX, "gt_mol_files", pid, pid + "_ligand.sdf"))
        out[pid + "_ca"], out[pid + "_lig"], out[pid + "_bonds"] = ca, lig, bonds
        d = np.linalg.norm(ca - lig.mean(0), axis=1)
        print(pid, "CA", len(ca), "ligand heavy atoms", len(lig), "bonds", bonds.shape[1] // 2, "pocket(20A)", int((d < 20).sum()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "real_geometry.npz"), ids=np.asarray(IDS), **out)


if __name__ == "__main__":
    sys.exit(main())
