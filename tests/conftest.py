import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import helpers as H  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(H.GOLDEN, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    return po.load_oracle()


@pytest.fixture(scope="session")
def genomes():
    return H.load_genomes()


@pytest.fixture(scope="session")
def reads2000():
    return H.make_reads(2000, seed=42)


@pytest.fixture(scope="session")
def toy_tax(oracle):
    c, p = H.toy_tax_arrays()
    return oracle.tax_from_pairs(c, p)


class DbCache:
    """Oracle-built databases of the 4-genome fixture, keyed by golden.json's db names."""

    def __init__(self, oracle, genomes, tax, golden):
        self.o, self.g, self.t, self.gold, self.dbs = oracle, genomes, tax, golden, {}

    def get(self, name):
        if name not in self.dbs:
            spec = self.gold["dbs"][name]
            db = self.o.db_new()
            for gi, taxid in enumerate(H.GENOME_TAXIDS):
                self.o.db_add_genome(db, self.t, H.genome_records(self.g, gi), taxid, spec["k"], spec["w"],
                                     spec["gaps"], spec["score"], spec["canon"], cast_mode=po.CAST_SATURATE)
            self.dbs[name] = db
        return self.dbs[name]


@pytest.fixture(scope="session")
def dbcache(oracle, genomes, toy_tax, golden):
    return DbCache(oracle, genomes, toy_tax, golden)
