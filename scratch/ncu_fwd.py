import sys
sys.path.insert(0, '.')
from scratch.ablate_tc import t
print(t(2052096, 128, 128, True, True, 0))
