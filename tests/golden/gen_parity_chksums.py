"""Golden checksums for bench.py's self-verification (`parity_check`) and the multi-GPU decomposition-invariance checks.

    python tests/golden/gen_parity_chksums.py        # writes tests/golden/parity_chksums.json

The CPU oracle advects the 1-degree tripolar case (360 x 300 x 50, 3 tracers, analytic flow scale cfl/12 so that no rank needs
a global reduction to generate its block) as ONE block and prints the reference's checksum -- mpp_chksum: the wrap-around int64
sum of the bit patterns over the compute domain (src/shared/mpp/include/mpp_chksum_int.h:20-38) -- of every th_tendency and
T_prog%wrk1.  Any decomposition of the same grid must reproduce them bit for bit (test/test_bit_reproducibility.py:17-64 of the
reference).  tests/test_oracle_invariants.py re-derives the file on the CPU, with a multi-block layout as well.
"""
import dataclasses
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def compute(layout=(1, 1)):
    from mom5_b200.synthetic import CASES, Generator
    from oracle.oracle import Oracle, split_blocks
    base = CASES["global_1deg"]
    spec = dataclasses.replace(base, ntr=3, flow_scale=base.cfl / 12.0)
    gb = Generator(spec).block(1, spec.ni, 1, spec.nj, ntr=3)
    dec = spec.decomposition(*layout)
    blocks = split_blocks(dec, gb)
    o = Oracle(dec, blocks)
    th = [[t.numpy().copy() for t in b.th_tendency] for b in blocks]
    ref = o.sweby_all_timed([[t.numpy() for t in b.T] for b in blocks], th, spec.dtime, nthreads=len(blocks))
    nb = len(blocks)
    return dict(th=[o.chksum([th[r][n] for r in range(nb)]) for n in range(3)],
                adv=[o.chksum([ref["adv"][r][n] for r in range(nb)]) for n in range(3)])


if __name__ == "__main__":
    out = dict(global_1deg_ntr3=dict(case="global_1deg", ni=360, nj=300, nk=50, ntr=3, flow_scale="cfl/12", **compute()))
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "parity_chksums.json")
    json.dump(out, open(p, "w"), indent=1)
    print(open(p).read())
