import ctypes as C, torch, sys
sys.path.insert(0, '.')
from scratch.quick_tc import run
