import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_sessionstart(session):
    # make sure the oracle (test infrastructure) is built; the product library must already
    # have been built by __graft_entry__.build()
    from oracle import oracle as orc
    orc.load()
