"""Pipeline timeline of CTA 0 of conv_tc_fwd_kernel from its clock64 stamps.
usage: python scratch/fwd_stamps.py N H W Cin Cout k [math]"""
import sys
import torch
sys.path.insert(0, '.')
from dl4ds_b200 import _lib
from dl4ds_b200._lib import MATH
N, H, W, Ci, Co, k = [int(v) for v in sys.argv[1:7]]
math = sys.argv[7] if len(sys.argv) > 7 else 'tf32x3'
dev = torch.device('cuda'); lib = _lib.load()
x = torch.randn(N, H, W, Ci, device=dev); w = torch.randn(k, k, Ci, Co, device=dev) * 0.05
y = torch.empty(N, H, W, Co, device=dev)
nb = lib.dl4ds_conv2d_fwd_workspace_bytes(N, H, W, Ci, H, W, Co, k, k, 1, 1, 1, MATH[math])
ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
dbg = torch.zeros(1100, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
for rep in range(3):
    dbg.zero_(); lib.dl4ds_debug_set_buffer(dbg.data_ptr())
    _lib.call('dl4ds_conv2d_fwd', x.data_ptr(), Ci, w.data_ptr(), None, None, 0, y.data_ptr(), Co, N, H, W, Ci, H, W, Co,
              k, k, 1, 1, k // 2, k // 2, 0, 0, 1, 0, MATH[math], ws.data_ptr(), st)
    torch.cuda.synchronize()
lib.dl4ds_debug_set_buffer(None)
t = dbg.cpu()[:1024].view(128, 8)
t0 = int(t[0][t[0] > 0].min())
names = ['tma:empty', 'mma:tile', 'mma:wait', 'mma:ready', 'mma:issued', 'spl:full', 'spl:done']
print('group ' + ' '.join('%10s' % n for n in names))
for g in range(128):
    if int(t[g].max()) == 0: break
    print('%5d ' % g + ' '.join('%10d' % (int(t[g, j]) - t0 if int(t[g, j]) else -1) for j in range(7)))
