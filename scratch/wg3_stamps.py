"""clock64 timeline of CTA 0 of conv_tc_wgrad3_kernel.  usage: python scratch/wg3_stamps.py N H W Ca Cb k"""
import sys
import torch
sys.path.insert(0, '.')
from dl4ds_b200 import _lib
from dl4ds_b200._lib import MATH
N, H, W, Ca, Cb, k = [int(v) for v in sys.argv[1:7]]
dev = torch.device('cuda')
lib = _lib.load()
P = torch.randn(N, H, W, Ca, device=dev); Q = torch.randn(N, H, W, Cb, device=dev)
dw = torch.zeros(k, k, Ca, Cb, device=dev)
dbg = torch.zeros(64 * 16 + 64, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
for rep in range(3):
    dbg.zero_()
    lib.dl4ds_debug_set_buffer(dbg.data_ptr())
    _lib.call('dl4ds_conv2d_wgrad', P.data_ptr(), Ca, Q.data_ptr(), Cb, dw.data_ptr(), N, H, W, Ca, H, W, Cb, k, k, 1, k // 2, k // 2, None, MATH['tf32x3'], st)
    torch.cuda.synchronize()
lib.dl4ds_debug_set_buffer(None)
t = dbg.cpu()[:1024].view(64, 16)
t0 = int(t[0, 12])
print('entry 0, prologue+wait done %d, scales known %d, accum ready %d, epilogue done %d' % (int(t[0, 13]) - t0, int(t[0, 14]) - t0, int(t[1, 12]) - t0, int(t[1, 13]) - t0))
print('chunk    P:start   P:fenced   M:full  M:issued')
for it in range(20):
    if int(t[it, 0]) == 0: break
    print('%5d %10d %10d %8d %9d' % (it, int(t[it, 0]) - t0, int(t[it, 1]) - t0, int(t[it, 4]) - t0, int(t[it, 5]) - t0))
