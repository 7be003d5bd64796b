import sys, numpy as np, torch
sys.path.insert(0, '.')
from dl4ds_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda')
def run(P, Q, KH, KW, pt, pl, math):
    N, H, W, Ca = P.shape; Cb = Q.shape[3]
    Pd = torch.from_numpy(P).to(dev); Qd = torch.from_numpy(Q).to(dev)
    dw = torch.zeros(KH, KW, Ca, Cb, device=dev)
    n0 = lib.dl4ds_tc_launch_count()
    _lib.call('dl4ds_conv2d_wgrad', Pd.data_ptr(), Ca, Qd.data_ptr(), Cb, dw.data_ptr(), N, H, W, Ca, H, W, Cb, KH, KW, 1, pt, pl, None, math, None)
    torch.cuda.synchronize()
    return dw.cpu().numpy(), lib.dl4ds_tc_launch_count() - n0
def ref(P, Q, KH, KW, pt, pl):
    N, H, W, Ca = P.shape; Cb = Q.shape[3]
    Pp = np.zeros((N, H + KH, W + KW, Ca), np.float64)
    Pp[:, pt:pt + H, pl:pl + W] = P
    dw = np.zeros((KH, KW, Ca, Cb))
    for kh in range(KH):
        for kw in range(KW):
            dw[kh, kw] = np.einsum('nyxa,nyxb->ab', Pp[:, kh:kh + H, kw:kw + W], Q.astype(np.float64))
    return dw
np.set_printoptions(linewidth=250, suppress=True)
# case 1: 1x1, Ca=8, Cb=8, one-hot P
H = W = 8
for (p0, a0) in [(0, 0), (1, 0), (0, 1), (9, 3), (63, 7)]:
    P = np.zeros((1, H, W, 8), np.float32); P.reshape(-1, 8)[p0, a0] = 1
    Q = (np.arange(64 * 8).reshape(1, H, W, 8)).astype(np.float32)
    d, n = run(P, Q, 1, 1, 0, 0, 2)
    print('onehot pix', p0, 'ca', a0, 'tc launches', n)
    print(d[0, 0])
    print('expected row', a0, '=', Q.reshape(-1, 8)[p0])
# case 2: random small ints
rng = np.random.default_rng(0)
for (Ca, Cb, KH, KW) in [(8, 8, 1, 1), (8, 8, 3, 3), (16, 16, 1, 1), (48, 192, 3, 3), (32, 64, 1, 1)]:
    P = rng.integers(-3, 4, (2, 8, 8, Ca)).astype(np.float32)
    Q = rng.integers(-3, 4, (2, 8, 8, Cb)).astype(np.float32)
    d, n = run(P, Q, KH, KW, KH // 2, KW // 2, 2)
    r = ref(P, Q, KH, KW, KH // 2, KW // 2)
    print('case', Ca, Cb, KH, KW, 'tc', n, 'max abs err', np.abs(d - r).max(), 'max ref', np.abs(r).max())
    if np.abs(d - r).max() > 0 and Ca <= 16:
        print('got\n', d[KH // 2, KW // 2][:8, :8], '\nref\n', r[KH // 2, KW // 2][:8, :8])
