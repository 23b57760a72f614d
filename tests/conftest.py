import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.dirname(os.path.abspath(__file__))
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)          # helper modules next to the tests (emulation.py)


# no network on any box: the LPIPS AlexNet trunk is the seeded stand-in the oracle uses too (loss/perceptual.py)
os.environ.setdefault("HFC_LPIPS_SYNTHETIC", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")
