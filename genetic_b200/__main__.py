"""python -m genetic_b200 paramfile.txt [-f]: the genetIC command line on the device (icgen.py)."""
import sys
import time

_t0 = time.perf_counter()
from .icgen import main  # noqa: E402

rc = main()
import os  # noqa: E402
if os.environ.get("GX_TIMING", "0") == "1":
    print("[timing] whole process after interpreter start %.3f s" % (time.perf_counter() - _t0))
sys.exit(rc)
