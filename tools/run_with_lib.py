#!/usr/bin/env python
"""developer helper: run a repo script against another build of the library
  python tools/run_with_lib.py build/libbrcnn_old.so bench.py --steps 2 ..."""
import os
import runpy
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
from boosting_rcnn_b200 import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[1])
script = sys.argv[2]
sys.argv = sys.argv[2:]
runpy.run_path(script, run_name='__main__')
