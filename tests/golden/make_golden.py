"""Regenerates the committed fixtures under tests/golden/ from the reference tree.

Run in the build container (needs /root/reference; the GPU box never runs this):
    python tests/golden/make_golden.py
* karate.csv            -- copy of the reference's only dataset fixture (datasets/karate.csv).
* reference_pins.json   -- literal expectations stated by the reference's own tests for this path
                           (file:line recorded next to each pin) so the tests do not have to read
                           /root/reference at run time.
"""
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def main():
    shutil.copyfile(os.path.join(REF, "datasets", "karate.csv"), os.path.join(HERE, "karate.csv"))
    pins = {
        "append_unique_docstring_example": {
            "source": "python/pylibwholegraph/pylibwholegraph/torch/graph_ops.py:22-27",
            "targets": [3, 11, 2, 10],
            "neighbors": [4, 5, 2, 11, 6, 9, 10, 5],
            "unique_prefix": [3, 11, 2, 10],
            "unique_tail_sorted": [4, 5, 6, 9],
        },
        "hetero_fanout_all": {
            "source": "python/cugraph-pyg/cugraph_pyg/tests/sampler/test_distributed_sampler.py:19-150",
            "srcs": [4, 5, 6, 7, 8, 9, 8, 9, 8, 7, 6, 5, 4, 5],
            "dsts": [0, 1, 2, 3, 3, 0, 4, 5, 6, 8, 7, 8, 9, 9],
            "eids": [0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5, 6, 7],
            "etps": [0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1],
            "vertex_type_offsets": [0, 4, 10],
            "seeds": [4, 5],
            "expect": {
                "etype0_hop0": {"eids": [0, 1], "srcs": [4, 5], "dsts": [0, 1]},
                "etype0_hop1": {"eids": [4, 5], "srcs": [8, 9], "dsts": [0, 3]},
                "etype1_hop0": {"eids": [5, 6, 7], "srcs": [4, 5, 5], "dsts": [8, 9, 9]},
                "etype1_hop1": {"eids": [0, 1, 2], "srcs": [8, 8, 9], "dsts": [4, 5, 6]},
            },
        },
        "gather_closed_form": {
            "source": "python/pylibwholegraph/pylibwholegraph/tests/wholegraph_torch/ops/test_wholegraph_gather_scatter.py:16-31",
            "rule": "table[i][d] = i + d ; gathered[k][d] == indices[k] + d",
        },
        "pcg32_published_kat": {
            "source": "pcg-c-basic pcg32-demo (initstate=42, initseq=54), published by the PCG authors",
            "initstate": 42,
            "initseq": 54,
            "outputs_hex": ["a15c02b7", "7b47f409", "ba1d3330", "83d2f293", "bfa4784b", "cbed606e"],
        },
        "sampler_launch_tables": {
            "source": "cpp/src/wholegraph_ops/unweighted_sample_without_replacement_func.cuh:445-447 and func_array :412-444",
            "warp_count": [1, 1, 1, 2, 2, 2, 4, 4, 4, 4, 4, 4] + [8] * 20,
            "items_per_thread": [1, 2, 3, 2, 3, 3, 2, 2, 3, 3, 3, 3, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4],
        },
    }
    with open(os.path.join(HERE, "reference_pins.json"), "w") as f:
        json.dump(pins, f, indent=1)


if __name__ == "__main__":
    main()
