import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def golden_cases():
    import json
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_transcripts.json")
    with open(path) as f:
        return json.load(f)["cases"]


@pytest.fixture(scope="session")
def golden_extreme_cases():
    import json
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_transcripts_extreme.json")
    with open(path) as f:
        return json.load(f)["cases"]
