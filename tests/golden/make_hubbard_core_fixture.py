"""Generates tests/golden/hubbard_doubles_core.json from two regression runs of the reference on the 12-site k-space
Hubbard chain (U = 1, 12 electrons) with `semi-stochastic doubles-core`: test_suite/mneci/cfqmc/hubbard_4_states (total
momentum 6: both open-shell electrons in the same eps = 0 orbital) and test_suite/mneci/kpfciqmc/hub_10_ft (total
momentum 0: one electron in each).  Recorded: lattice and electron number from the input, and what the reference
printed -- reference energy, size of the deterministic space, its lowest eigenvalue relative to the reference
(`Deterministic subspace correlation energy`).  Run in the build container (/root/reference present)."""
import glob
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"


def case(rel):
    d = os.path.join(REF, "test_suite", rel)
    inp = open(os.path.join(d, "neci.inp")).read()
    bench = open(glob.glob(os.path.join(d, "benchmark*"))[0]).read()
    cell = [int(x) for x in re.search(r"cell\s+(\d+)\s+(\d+)\s+(\d+)", inp).groups()]
    return dict(source="test_suite/" + rel, cell=cell, u=float(re.search(r"\n\s*u\s+(\S+)", inp).group(1)),
                b=float(re.search(r"\n\s*b\s+(\S+)", inp).group(1)), electrons=int(re.search(r"electrons\s+(\d+)", inp).group(1)),
                sym=[int(x) for x in re.search(r"sym\s+(\d+)\s+(\d+)\s+(\d+)\s+(\d+)", inp).groups()],
                reference_energy=float(re.search(r"Reference Energy set to:\s+(-?[\d.]+)", bench).group(1)),
                core_size=int(re.search(r"Total size of deterministic space:\s+(\d+)", bench).group(1)),
                core_correlation_energy=float(re.search(r"Deterministic subspace correlation energy:\s+(-?[\d.]+)", bench).group(1)))


def main():
    out = dict(total_momentum_6=case("mneci/cfqmc/hubbard_4_states"), total_momentum_0=case("mneci/kpfciqmc/hub_10_ft"))
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hubbard_doubles_core.json")
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst, out)


if __name__ == "__main__":
    main()
