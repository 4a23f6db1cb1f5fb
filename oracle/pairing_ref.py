"""TEST INFRASTRUCTURE -- the big-integer BLS12-377 pairing model (tools/pairing_model.py), under the name the tests import.
The product's host pairing (csrc/pairing.h) is compared with it bit for bit (tests/test_verifier.py)."""
from tools.pairing_model import *  # noqa: F401,F403
from tools.pairing_model import FE, G1, G2, ORDER, X, e1mul, e2mul, f12pow, F12ONE, pairing, q, r  # noqa: F401
