"""Generates tests/golden/beh_open_shell.npz from the reference's regression case
test_suite/neci/rdm_singlerun/parallel/BeH_open_shell_explicit: the FCIDUMP as it is (19 orbitals, 5 electrons, Ms = -1/2
in the run, C2v labels) and what the reference printed after `freeze 2 0` and `semi-stochastic doubles-core`:
reference determinant and energy, size of the deterministic space, its lowest eigenvalue relative to the reference.
Run in the build container (/root/reference present); the tests read only the .npz."""
import glob
import os
import re
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from neci_stable_b200 import fcidump  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
CASE = os.path.join(REF, "test_suite", "neci", "rdm_singlerun", "parallel", "BeH_open_shell_explicit")


def main():
    d = fcidump.read_fcidump(os.path.join(CASE, "FCIDUMP"))
    inp = open(os.path.join(CASE, "neci.inp")).read()
    bench = open(glob.glob(os.path.join(CASE, "benchmark*"))[0]).read()
    ref_det = [int(x) for x in re.search(r"Generated reference determinants:\s*\n\(\s*([\d,\s]+)\)", bench).group(1).replace(",", " ").split()]
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "beh_open_shell.npz")
    np.savez_compressed(
        dst, norb=d.norb, nelec=d.nelec, orbsym=np.array(d.orbsym), ecore=d.ecore, eps=np.array(d.eps),
        h1=np.array(d.h1), eri=np.array(d.eri),
        input_electrons=int(re.search(r"electrons\s+(\d+)", inp).group(1)),
        input_spin_restrict=int(re.search(r"spin-restrict\s+(-?\d+)", inp).group(1)),
        input_freeze=np.array([int(x) for x in re.search(r"freeze\s+(\d+)\s+(\d+)", inp).groups()]),
        reference_det=np.array(ref_det),
        reference_energy=float(re.search(r"Reference Energy set to:\s+(-?[\d.]+)", bench).group(1)),
        core_size=int(re.search(r"Total size of deterministic space:\s+(\d+)", bench).group(1)),
        core_correlation_energy=float(re.search(r"Deterministic subspace correlation energy:\s+(-?[\d.]+)", bench).group(1)))
    print("wrote", dst, os.path.getsize(dst), "bytes", d.norb, d.nelec, len(d.h1), len(d.eri), ref_det)


if __name__ == "__main__":
    main()
