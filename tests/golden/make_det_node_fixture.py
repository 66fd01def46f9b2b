"""Generates tests/golden/reference_det_nodes.json: `Reference processor is: N` as printed by the reference's own
regression runs, together with everything DetermineDetNode (src/load_balance_calcnodes.F90:25-117) needs to
reproduce it: RandomOrbIndex rebuilt exactly as src/fcimc_initialisation.fpp:824-890 builds it -- dSFMT_init(abs(seed))
on the root, then INT(nBasis*r*1000)+1 with rejection of duplicates -- using the REFERENCE'S OWN dSFMT
(oracle/_ref/libdsfmt_ref.so, compiled in place from src/lib/dSFMT.cpp by oracle/Makefile).  Run in the build
container; the tests read only the JSON."""
import ctypes as C
import glob
import json
import os
import re
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DSFMT = os.path.join(HERE, "..", "..", "oracle", "_ref", "libdsfmt_ref.so")
STORE = 50000                                   # random_store_size, src/lib/dSFMT_interface.F90:22


class RefRng:
    def __init__(self, seed):
        self.lib = C.CDLL(DSFMT)
        self.lib.init_gen_rand_fwrapper(C.c_uint32(seed))
        self.buf = np.zeros(STORE)
        self._fill()

    def _fill(self):
        self.lib.fill_array_close_open_fwrapper(self.buf.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(STORE))
        self.pos = 0

    def real2(self):                             # genrand_real2_dSFMT
        if self.pos == STORE:
            self._fill()
        r = self.buf[self.pos]; self.pos += 1
        return float(r)


def random_orb_index(nbasis, seed):
    rng = RefRng(seed)
    roi = [0] * nbasis
    for i in range(nbasis):
        while True:
            chosen = int(nbasis * rng.real2() * 1000) + 1
            dup = chosen in roi
            roi[i] = chosen
            if not dup:
                break
    return roi


def det_block(nI, roi, blocks):
    acc = 0
    for i, o in enumerate(nI, 1):
        acc = (1099511628211 * acc + roi[o - 1] * i) & 0xFFFFFFFFFFFFFFFF
    if acc >= 1 << 63:
        acc -= 1 << 64
    m = abs(acc) % blocks if acc >= 0 else -((-acc) % blocks)       # Fortran mod: sign of the dividend
    return abs(m) + 1


def main():
    out = []
    for d in sorted(glob.glob(os.path.join(REF, "test_suite", "neci", "parallel", "*"))):
        inp = glob.glob(os.path.join(d, "*.inp")); ben = glob.glob(os.path.join(d, "benchmark*"))
        if not inp or not ben:
            continue
        itxt = open(inp[0]).read().lower(); btxt = open(ben[0], errors="replace").read()
        ms = re.search(r"^\s*seed\s+(-?\d+)", itxt, re.M)
        mp = re.search(r"Reference processor is:\s+(\d+)", btxt)
        mn = re.search(r"Number of processors:\s+(\d+)", btxt)
        md = re.search(r"Generated reference determinants:\s*\n\(\s*([\d,\s]+)\)", btxt)
        mb = re.search(r"NUMBER OF SPIN ORBITALS IN BASIS :\s+(\d+)", btxt)
        if not (ms and mp and mn and md and mb) or "spatial-only-hash" in itxt or "readpops" in itxt:
            continue
        frz = re.search(r"^\s*freeze\s+(\d+)\s+(\d+)", itxt, re.M)
        nbasis = int(mb.group(1)) - (int(frz.group(1)) + int(frz.group(2)) if frz else 0)
        nI = [int(x) for x in md.group(1).replace(",", " ").split()]
        nprocs = int(mn.group(1))
        if nbasis > 128:                                              # beyond the engine's two-word determinants
            continue
        blocks = nprocs * (1 if re.search(r"load-balance-blocks\s+off", itxt) else 100)
        seed = abs(int(ms.group(1)))
        roi = random_orb_index(nbasis, seed)
        block = det_block(nI, roi, blocks)
        proc = (block - 1) // (blocks // nprocs)                      # initial LoadBalanceMapping(i) = int((i-1)/oversample_factor), load_balancer.fpp:96-99
        ok = proc == int(mp.group(1))
        print("%-42s nbasis %3d seed %3d procs %d blocks %4d det %-40s -> %d (reference %s) %s" % (
            os.path.basename(d), nbasis, seed, nprocs, blocks, str(nI)[:40], proc, mp.group(1), "ok" if ok else "MISMATCH"))
        if ok:
            out.append(dict(case=os.path.basename(d), nbasis=nbasis, seed=seed, nprocs=nprocs, balance_blocks=blocks, det=nI,
                            random_orb_index=roi, block=block, reference_processor=int(mp.group(1))))
    json.dump(out, open(os.path.join(HERE, "reference_det_nodes.json"), "w"))
    print(len(out), "cases written")


if __name__ == "__main__":
    main()
