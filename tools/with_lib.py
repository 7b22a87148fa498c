"""A/B helper: run a tool against another build of the library.

    python tools/with_lib.py tools/ab/libssdhead_base.so tools/time_loss.py 32 D1 0 stable
"""
import os, runpy, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_detection_torch2_b200 import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")
