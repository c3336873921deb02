"""Golden vectors for the random stream geometry of the samplers, from an implementation this project did not write:
M. O'Neill's canonical pcg-cpp (`pcg32` = setseq_xsh_rr_64_32), as vendored by Apache Arrow in the pyarrow wheel
(pyarrow/include/arrow/vendored/pcg/pcg_random.hpp).

RAFT's `PCGenerator(seed, subsequence, offset)` (raft/random/detail/rng_device.cuh, RAFT 26.10, un-vendored dependency of the
reference: cpp/src/wholegraph_ops/raft_random_gen.cu:32-53, unweighted_sample_without_replacement_func.cuh:137) is that
generator: state = 0; inc = 2*subsequence + 1; step; state += seed; step; skipahead(offset) -- i.e. pcg32(seed, subsequence)
followed by advance(offset); `PCGenerator(DeviceState{seed, base}, subsequence)` passes (seed, base + subsequence, subsequence).
This script pins that chain -- seeding, stream selection and the O(log n) advance -- for a spread of (seed, subsequence)
pairs, including subsequences beyond 2^32 (call groups index rows as 32 * row + lane).

    python tests/golden/make_pcg_vectors.py      -> tests/golden/pcg_canonical_vectors.json
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = r'''
#include <cstdint>
#include <cstdio>
#include "arrow/vendored/pcg/pcg_random.hpp"
int main(int argc, char** argv)
{
  // (seed, subsequence) pairs on stdin; for each: pcg32(seed, subsequence); advance(subsequence); eight draws
  unsigned long long seed, sub;
  while (std::scanf("%llu %llu", &seed, &sub) == 2) {
    ::arrow_vendored::pcg32 rng((uint64_t)seed, (uint64_t)sub);
    rng.advance((uint64_t)sub);
    for (int i = 0; i < 8; i++)
      std::printf("%u%c", (unsigned)rng(), i == 7 ? '\n' : ' ');
  }
  return 0;
}
'''


def canonical(pairs):
    import pyarrow

    inc = os.path.join(os.path.dirname(pyarrow.__file__), "include")
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "pcgvec.cpp"), os.path.join(d, "pcgvec")
        open(src, "w").write(SRC)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", inc, src, "-o", exe])
        out = subprocess.run([exe], input="".join("%d %d\n" % p for p in pairs), capture_output=True, text=True, check=True).stdout
    return [[int(x) for x in line.split()] for line in out.strip().splitlines()]


PAIRS = [(0, 0), (0, 1), (62, 0), (62, 1), (62, 31), (62, 32), (62, 33), (62, 1023), (62, 32 * 1024 + 7), (12345678901234567, 987654321),
         (62 + 0x9E3779B97F4A7C15 & 0xFFFFFFFFFFFFFFFF, 32 * 2_500_000 + 31), (1, (1 << 32) + 5), (0xFFFFFFFFFFFFFFFF, (1 << 40) + 12345),
         (42, (1 << 62) + 3)]

if __name__ == "__main__":
    vec = canonical(PAIRS)
    json.dump({"source": "pcg-cpp pcg32 (arrow_vendored, pyarrow %s): pcg32(seed, subsequence); advance(subsequence); 8 draws" % __import__("pyarrow").__version__,
               "pairs": [list(p) for p in PAIRS], "draws": vec}, open(os.path.join(HERE, "pcg_canonical_vectors.json"), "w"), indent=1)
    print("wrote", len(vec), "vectors")
