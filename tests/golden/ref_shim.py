"""The reference import shim moved to ``oracle/ref_shim.py`` (the bench's reference arm uses it too); re-exported here
for ``tests/golden/gen_golden.py`` / ``gen_kat.py``."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.ref_shim import *  # noqa: F401,F403,E402
from oracle.ref_shim import install, raw  # noqa: F401,E402
