"""python -m genetic_b200 paramfile.txt [-f]: the genetIC command line on the device (icgen.py)."""
import sys

from .icgen import main

sys.exit(main())
