"""Total clocks of CTA 0 (entry -> epilogue done) of conv_tc_wgrad2 (DL4DS_WGRAD3=0) or conv_tc_wgrad3.
usage: [DL4DS_WGRAD3=0] python scratch/wgrad_compare.py"""
import os, sys
import torch
sys.path.insert(0, '.')
from dl4ds_b200 import _lib
from dl4ds_b200._lib import MATH
lib = _lib.load()
dev = torch.device('cuda')
st = torch.cuda.current_stream().cuda_stream
v3 = os.environ.get('DL4DS_WGRAD3', '1') != '0'
for (N, H, W, Ca, Cb, k) in [(64, 32, 32, 16, 16, 3), (64, 32, 32, 24, 24, 3), (64, 32, 32, 32, 32, 3), (64, 32, 32, 40, 40, 3),
                              (64, 32, 32, 48, 48, 3), (64, 32, 32, 8, 16, 3), (64, 32, 32, 40, 48, 3), (64, 64, 64, 48, 32, 3)]:
    P = torch.randn(N, H, W, Ca, device=dev); Q = torch.randn(N, H, W, Cb, device=dev)
    dw = torch.zeros(k, k, Ca, Cb, device=dev)
    dbg = torch.zeros(64 * 16 + 64, dtype=torch.int64, device=dev)
    for rep in range(3):
        dbg.zero_()
        lib.dl4ds_debug_set_buffer(dbg.data_ptr())
        _lib.call('dl4ds_conv2d_wgrad', P.data_ptr(), Ca, Q.data_ptr(), Cb, dw.data_ptr(), N, H, W, Ca, H, W, Cb, k, k, 1, k // 2, k // 2, None, MATH['tf32x3'], st)
        torch.cuda.synchronize()
    lib.dl4ds_debug_set_buffer(None)
    d = dbg.cpu()
    if v3:
        t = d[:1024].view(64, 16)
        print('wgrad3 %2d->%2d @%dx%d: total %6d clk (pass1 %d, epilogue %d)' % (Ca, Cb, H, W, int(t[1, 13] - t[0, 12]), int(t[0, 14] - t[0, 13]), int(t[1, 13] - t[1, 12])))
    else:
        kk = d[1024:1028]
        print('wgrad2 %2d->%2d @%dx%d: total %6d clk (epilogue %d)' % (Ca, Cb, H, W, int(kk[3] - kk[0]), int(kk[3] - kk[2])))
