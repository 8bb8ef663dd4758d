import json
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def kat():
    with open(os.path.join(GOLDEN, "kernel_kat.json")) as fh:
        return json.load(fh)


def case_newick(spec):
    """Newick text of one side of a KAT case."""
    if "file" in spec:
        from oracle import phylok_oracle as po
        trees = po._split_trees_text(open(os.path.join(GOLDEN, "trees", spec["file"])).read())
        return trees[spec.get("index", 0)]
    return spec["newick"]


PREP = {"raw": (0, "none"), "kamphir:mean": (3, "mean"), "kamphir:median": (3, "median"), "ladder:none": (2, "none")}
