import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_gpu_kernels import _wgrad_stream
for (M, Lq, Cc, N) in [(128 * 16032, 128, 128, 128), (128 * 4000, 128, 128, 128), (128 * 16032, 128, 64, 64), (128*16032, 128, 192, 64)]:
    torch.manual_seed(M + Cc)
    for bn in (False, True):
        for rep in range(2):
            print((M, Lq, Cc, N), bn, "e_w %.3e e_b %.3e" % _wgrad_stream(M, Lq, Cc, N, bn), flush=True)
