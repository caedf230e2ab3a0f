import sys
sys.path.insert(0, '.')
from blangsdk_b200.engine import microbench
for c in (1, 2, 4):
    print('chains', c, 'cycles per DFMA (per warp, 1 warp/SMSP)', microbench(10 + c))
