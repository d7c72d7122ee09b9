"""A/B helper: run bench.py with another build of libvlk.so (same box, same process layout): 
   python scripts/probe/bench_with_lib.py scripts/probe/libvlk_x.so --workload pretrain --steps 4 ..."""
import os, runpy, sys
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, root)
from gpt2_vision_language_b200 import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = [os.path.join(root, "bench.py")] + sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")
