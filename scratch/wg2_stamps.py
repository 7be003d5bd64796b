"""Pipeline timeline of CTA (0,0) of conv_tc_wgrad2_kernel from its clock64 stamps.
usage: python scratch/wg2_stamps.py N H W Ca Cb k [math]"""
import sys
import torch
sys.path.insert(0, '.')
from dl4ds_b200 import _lib
from dl4ds_b200._lib import MATH

N, H, W, Ca, Cb, k = [int(v) for v in sys.argv[1:7]]
math = sys.argv[7] if len(sys.argv) > 7 else 'tf32x3'
dev = torch.device('cuda')
lib = _lib.load()
P = torch.randn(N, H, W, Ca, device=dev)
Q = torch.randn(N, H, W, Cb, device=dev)
dw = torch.zeros(k, k, Ca, Cb, device=dev)
dbg = torch.zeros(64 * 16 + 64, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
names = ['tma:rfree', 'mma:qfull', 'mma:afull0', 'mma:issued', 'q:rfull', 'q:qempty', 'q:done', 'q:arrived',
         'a:b0-go', 'a:b0-done', 'a:b0-arr', 'a:rfree']
for rep in range(3):
    dbg.zero_()
    lib.dl4ds_debug_set_buffer(dbg.data_ptr())
    _lib.call('dl4ds_conv2d_wgrad', P.data_ptr(), Ca, Q.data_ptr(), Cb, dw.data_ptr(), N, H, W, Ca, H, W, Cb, k, k, 1,
              k // 2, k // 2, None, MATH[math], st)
    torch.cuda.synchronize()
lib.dl4ds_debug_set_buffer(None)
k = dbg.cpu()[1024:1028]
print('kernel stamps (clk): setup %d, main loop %d, epilogue %d' % (int(k[1] - k[0]), int(k[2] - k[1]), int(k[3] - k[2])))
t = dbg.cpu()[:1024].view(64, 16)
t0 = int(t[0][t[0] > 0].min())
print('chunk ' + ' '.join('%10s' % n for n in names))
for it in range(64):
    if int(t[it].max()) == 0:
        break
    print('%5d ' % it + ' '.join('%10d' % (int(t[it, j]) - t0 if int(t[it, j]) else -1) for j in range(12)))
