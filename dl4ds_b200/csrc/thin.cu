// HBM-bound kernels for the NARROW layers of the DL4DS graphs (the HR tail: 8 / 1 channels at
// 128 x 128 -- sp_postups.py:205-212 -- and the 1-channel stem, sp_postups.py:134).  These layers
// cannot feed a 128 x N tensor-core tile (Cout 1 / 8) and are bandwidth-bound anyway, so they get
// CUDA-core kernels built around data reuse in shared memory and registers:
//
//   thin_wgrad_kernel<CA,CB,KS>  3x3 (and 7x7: one kernel row per thread) stride-1 weight gradient for Ca, Cb in {1, 8}: an (rows x cols)
//       tile of P (with halo) and Q lives in shared memory; a thread owns a TA x TB block of
//       (ca, cb) pairs for ALL nine taps (9*TA*TB accumulators) and walks 32 pixels of one row with
//       a 3x3 sliding register window, i.e. 4 shared loads per 72 FMAs for Ca = Cb = 8.
//   bias_act_bwd_vec4_kernel  dz = dy * act'(y) (+ space_to_depth un-shuffle), dbias += sum dz with
//       16-byte accesses.
#include "common.cuh"

namespace dl4ds {

// -------------------------------------------------------------------------------------------------
// thin 3x3 wgrad
// -------------------------------------------------------------------------------------------------
template <int CA, int CB, int KS>
struct ThinCfg {
    // 3x3: a thread owns a 2 x 4 block of (ca, cb) pairs for all nine taps.  7x7 (the ConvNeXt stem / tail, never
    // 8 x 8 channels): a thread owns ONE kernel row, all seven taps of it and every (ca, cb) pair.
    static constexpr int TA = KS == 3 ? (CA >= 2 ? 2 : 1) : CA;
    static constexpr int TB = KS == 3 ? (CB >= 4 ? 4 : 1) : CB;
    static constexpr int KHT = KS == 3 ? 3 : 1;             // kernel rows per thread
    static constexpr int KG = KS / KHT;                     // kernel-row groups
    static constexpr int CPS = (CA / TA) * (CB / TB);       // channel-block threads per kernel-row group
    static constexpr int TPS = CPS * KG;                    // threads per pixel stream
    static constexpr int STREAMS = 256 / TPS;
    static constexpr int SEG = 32;                          // pixels per stream per tile: the largest choice (the launcher may pick 16 / 8)
};

struct ThinWgradArgs {
    const float* P; const float* Q; float* dw;
    int p_ld, q_ld;
    int N, H, W, pad_t, pad_l;
    int TW, TH, tiles_x, tiles_y, ntiles;
    int p_pitch, q_pitch;       // shared-memory row pitches in floats
    int seg;                    // pixels a stream walks per tile
};

template <int N>
__device__ __forceinline__ void lds_vec(const float* src, float (&out)[N]) {
    if constexpr (N % 4 == 0) {
#pragma unroll
        for (int i = 0; i < N; i += 4) {
            const float4 v = *reinterpret_cast<const float4*>(src + i);
            out[i] = v.x; out[i + 1] = v.y; out[i + 2] = v.z; out[i + 3] = v.w;
        }
    } else if constexpr (N == 2) {
        const float2 v = *reinterpret_cast<const float2*>(src);
        out[0] = v.x; out[1] = v.y;
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = src[i];
    }
}

template <int CA, int CB, int KS>
__global__ void __launch_bounds__(256) thin_wgrad_kernel(ThinWgradArgs a) {
    pdl_launch_dependents();    // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    a.P = pdl_after_wait(a.P);
    a.Q = pdl_after_wait(a.Q);
    using C = ThinCfg<CA, CB, KS>;
    constexpr int TA = C::TA, TB = C::TB, TPS = C::TPS, KHT = C::KHT, HALO = KS - 1;
    const int SEG = a.seg;
    extern __shared__ float sm[];
    float* Ps = sm;                                          // (TH+HALO) rows x p_pitch
    float* Qs = sm + (((size_t)(a.TH + HALO) * a.p_pitch + 3) & ~(size_t)3);   // TH rows x q_pitch (16-byte aligned)

    const int tid = threadIdx.x;
    const int stream = tid / TPS, sub = tid % TPS;
    const int kh0 = (sub / C::CPS) * KHT, csub = sub % C::CPS;
    const int ca0 = (csub / (CB / TB)) * TA, cb0 = (csub % (CB / TB)) * TB;
    const int segs = a.TW / SEG;
    // consecutive streams = consecutive rows (row pitches are chosen so that they hit distinct banks)
    const int srow = stream % a.TH, sseg = stream / a.TH;
    const bool active = sseg < segs;

    float acc[KHT][KS][TA][TB];
#pragma unroll
    for (int i = 0; i < KHT; ++i)
#pragma unroll
        for (int j = 0; j < KS; ++j)
#pragma unroll
            for (int u = 0; u < TA; ++u)
#pragma unroll
                for (int v = 0; v < TB; ++v) acc[i][j][u][v] = 0.0f;

    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int img = tile / (a.tiles_x * a.tiles_y);
        const int trem = tile - img * a.tiles_x * a.tiles_y;
        const int ty = trem / a.tiles_x, tx = trem - ty * a.tiles_x;
        const int y0 = ty * a.TH, x0 = tx * a.TW;
        __syncthreads();       // previous tile fully consumed
        // ---- P tile with halo: rows y0-pad_t .. +TH+HALO, cols x0-pad_l .. +TW+HALO, CA channels
        {
            const int cols = a.TW + HALO, rows = a.TH + HALO;
            if constexpr (CA % 4 == 0) {
                const int v4 = CA / 4, total = rows * cols * v4;
                for (int i = tid; i < total; i += 256) {
                    const int c4 = i % v4, px = (i / v4) % cols, r = i / (v4 * cols);
                    const int gy = y0 + r - a.pad_t, gx = x0 + px - a.pad_l;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W)
                        v = __ldg(reinterpret_cast<const float4*>(a.P + ((int64_t)(img * a.H + gy) * a.W + gx) * a.p_ld) + c4);
                    *reinterpret_cast<float4*>(Ps + (size_t)r * a.p_pitch + px * CA + c4 * 4) = v;
                }
            } else {
                const int total = rows * cols * CA;
                for (int i = tid; i < total; i += 256) {
                    const int c = i % CA, px = (i / CA) % cols, r = i / (CA * cols);
                    const int gy = y0 + r - a.pad_t, gx = x0 + px - a.pad_l;
                    float v = 0.f;
                    if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W)
                        v = __ldg(a.P + ((int64_t)(img * a.H + gy) * a.W + gx) * a.p_ld + c);
                    Ps[(size_t)r * a.p_pitch + px * CA + c] = v;
                }
            }
        }
        // ---- Q tile
        if constexpr (CB % 4 == 0) {
            const int v4 = CB / 4, total = a.TH * a.TW * v4;
            for (int i = tid; i < total; i += 256) {
                const int c4 = i % v4, px = (i / v4) % a.TW, r = i / (v4 * a.TW);
                const float4 v = __ldg(reinterpret_cast<const float4*>(
                                           a.Q + ((int64_t)(img * a.H + y0 + r) * a.W + x0 + px) * a.q_ld) + c4);
                *reinterpret_cast<float4*>(Qs + (size_t)r * a.q_pitch + px * CB + c4 * 4) = v;
            }
        } else {
            const int total = a.TH * a.TW * CB;
            for (int i = tid; i < total; i += 256) {
                const int c = i % CB, px = (i / CB) % a.TW, r = i / (CB * a.TW);
                Qs[(size_t)r * a.q_pitch + px * CB + c] =
                    __ldg(a.Q + ((int64_t)(img * a.H + y0 + r) * a.W + x0 + px) * a.q_ld + c);
            }
        }
        __syncthreads();
        if (active) {
            // window win[kh][kw][ta] = P at (row srow+kh0+kh, col x+kw) in halo coordinates
            float win[KHT][KS][TA];
            const int xs = sseg * SEG;
            const float* prow[KHT];
#pragma unroll
            for (int kh = 0; kh < KHT; ++kh) prow[kh] = Ps + (size_t)(srow + kh0 + kh) * a.p_pitch + ca0;
            const float* qrow = Qs + (size_t)srow * a.q_pitch + cb0;
#pragma unroll
            for (int kh = 0; kh < KHT; ++kh)
#pragma unroll
                for (int kw = 1; kw < KS; ++kw) lds_vec<TA>(prow[kh] + (xs + kw - 1) * CA, win[kh][kw]);
#pragma unroll 4
            for (int x = xs; x < xs + SEG; ++x) {
#pragma unroll
                for (int kh = 0; kh < KHT; ++kh) {
#pragma unroll
                    for (int kw = 0; kw < KS - 1; ++kw)
#pragma unroll
                        for (int u = 0; u < TA; ++u) win[kh][kw][u] = win[kh][kw + 1][u];
                    lds_vec<TA>(prow[kh] + (x + KS - 1) * CA, win[kh][KS - 1]);
                }
                float q[TB];
                lds_vec<TB>(qrow + x * CB, q);
                // (packed FFMA2 was tried here: 74 -> 102 us, the register pairs cost more moves than they save)
#pragma unroll
                for (int kh = 0; kh < KHT; ++kh)
#pragma unroll
                    for (int kw = 0; kw < KS; ++kw)
#pragma unroll
                        for (int u = 0; u < TA; ++u)
#pragma unroll
                            for (int v = 0; v < TB; ++v)
                                acc[kh][kw][u][v] = fmaf(win[kh][kw][u], q[v], acc[kh][kw][u][v]);
            }
        }
    }
    // ---- block reduction, then one global atomic per weight.  Every thread parks its accumulators in shared memory
    // ([element][stream]; the tiles are dead by now) and a warp per element sums the streams.  (Shared-memory atomics
    // here cost more than the whole main loop: all streams hit the same few addresses, 16- to 32-way conflicts per
    // instruction -- r02cfg5h: the kernel got SLOWER with more threads active.)
    constexpr int NACC = KHT * KS * TA * TB, NEL = KS * KS * CA * CB, SP = C::STREAMS;
    static_assert(NEL == TPS * NACC, "every weight element is owned by exactly one (sub, accumulator) pair");
    __syncthreads();
    if (stream < C::STREAMS) {
        float* dst = sm + (size_t)sub * NACC * SP + stream;
#pragma unroll
        for (int kh = 0; kh < KHT; ++kh)
#pragma unroll
            for (int kw = 0; kw < KS; ++kw)
#pragma unroll
                for (int u = 0; u < TA; ++u)
#pragma unroll
                    for (int v = 0; v < TB; ++v) dst[(size_t)(((kh * KS + kw) * TA + u) * TB + v) * SP] = acc[kh][kw][u][v];
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    for (int e = warp; e < NEL; e += 8) {
        float v = 0.f;
        for (int st = lane; st < SP; st += 32) v += sm[(size_t)e * SP + st];
        v = warp_sum(v);
        if (lane == 0) {
            const int esub = e / NACC, j = e - esub * NACC;
            const int ekh0 = (esub / C::CPS) * KHT, ecs = esub % C::CPS;
            const int eca0 = (ecs / (CB / TB)) * TA, ecb0 = (ecs % (CB / TB)) * TB;
            const int vv = j % TB, uu = (j / TB) % TA, kw = (j / (TB * TA)) % KS, kh = j / (TB * TA * KS);
            atomicAdd(a.dw + (((ekh0 + kh) * KS + kw) * CA + eca0 + uu) * CB + ecb0 + vv, v);
        }
    }
}

template <int CA, int CB, int KS>
static int launch_thin(ThinWgradArgs a, cudaStream_t st) {
    using C = ThinCfg<CA, CB, KS>;
    constexpr int HALO = KS - 1;
    // shared-memory budget: shrink the tile height until both tiles fit in ~96 KB
    auto bytes = [&](int t) {
        const int pp = ((a.TW + HALO) * CA + 31) / 32 * 32 + CA, qp = (a.TW * CB + 31) / 32 * 32 + CB;
        return (size_t)((((t + HALO) * pp + 3) & ~3) + t * qp) * 4;
    };
    // Pixels per stream (32 / 16 / 8) and tile height: a stream is a serial walk along a row segment (shared-memory
    // latency per pixel), so what matters is how many streams run at once -- every thread of the block busy (tile
    // rows x segments = streams) and at least two blocks per SM.  ncu (r02_thinw): 32-pixel segments of 16-row tiles
    // left half of the <1,8> block idle (three quarters of <1,1>) on a 128-block grid at 4 x 256 x 256 -- 65 us at 12 %
    // issue utilisation.  Of the choices that fill the block the longest segment wins (least window warm-up).
    int th = 1, best_seg = C::SEG;
    long best_score = -1;
    for (int seg = C::SEG; seg >= 8; seg >>= 1) {
        if (a.TW % seg) continue;
        const int segs = a.TW / seg;
        int t = C::STREAMS / segs;
        if (t < 1) continue;
        if (t > 16) t = 16;        // >= 512 tiles for the 1-channel cases (r01b: 128 blocks of 64-row tiles, 74 us)
        if (t > a.H) t = a.H;
        while (t > 1 && a.H % t) --t;
        while (t > 1 && bytes(t) > 96 * 1024) { --t; while (t > 1 && a.H % t) --t; }
        const long ntl = (long)a.N * (a.W / a.TW) * (a.H / t);
        const long blocks = ntl < 4L * kNumSMs ? ntl : 4L * kNumSMs;
        const long score = blocks * t * segs;             // streams in flight across the grid
        if (best_score < 0 || score * 5 > best_score * 6) { best_score = score; best_seg = seg; th = t; }    // shorter only for > 20 % more
    }
    a.seg = best_seg;
    a.TH = th;
    a.tiles_x = a.W / a.TW;
    a.tiles_y = a.H / th;
    a.ntiles = a.N * a.tiles_x * a.tiles_y;
    // row pitch = CA (mod 32) floats: consecutive rows land on consecutive bank groups
    a.p_pitch = ((a.TW + HALO) * CA + 31) / 32 * 32 + CA;
    a.q_pitch = (a.TW * CB + 31) / 32 * 32 + CB;
    size_t smem = bytes(th);
    const size_t red_bytes = (size_t)KS * KS * CA * CB * C::STREAMS * 4;      // the final [element][stream] reduction buffer
    if (smem < red_bytes) smem = red_bytes;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(thin_wgrad_kernel<CA, CB, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        attr = true;
    }
    int blocks_per_sm = (int)((200 * 1024) / (smem + 4096));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    if (blocks_per_sm > 4) blocks_per_sm = 4;
    int grid = kNumSMs * blocks_per_sm;
    if (grid > a.ntiles) grid = a.ntiles;
    launch_pdl(8, thin_wgrad_kernel<CA, CB, KS>, dim3(grid), dim3(256), smem, st, a);
    return check_launch("thin_wgrad_kernel");
}

// DL4DS_E_UNSUPPORTED when the shape is outside this kernel's domain
int conv2d_wgrad_thin(const WgradArgs& w, cudaStream_t st) {
    const bool k3 = w.KH == 3 && w.KW == 3, k7 = w.KH == 7 && w.KW == 7;
    if (!(k3 || k7) || w.stride != 1 || w.Hp != w.Hq || w.Wp != w.Wq) return DL4DS_E_UNSUPPORTED;
    const bool c1w = k3 && w.Ca == 1 && (w.Cb == 16 || w.Cb == 32 || w.Cb == 48 || w.Cb == 64);     // ConvBlock_aux/conv1 (cfg3: 1 -> 48)
    if (c1w) {
        if (w.stride != 1 || w.Hp != w.Hq || w.Wp != w.Wq || w.Wq % 32 || (w.Wq > 128 && w.Wq % 128) || w.NQ < 16384 ||
            w.q_ld % 4 || (reinterpret_cast<uintptr_t>(w.Q) & 15))
            return DL4DS_E_UNSUPPORTED;
        ThinWgradArgs a;
        a.P = w.P; a.Q = w.Q; a.dw = w.dw; a.p_ld = w.p_ld; a.q_ld = w.q_ld;
        a.N = w.N; a.H = w.Hq; a.W = w.Wq; a.pad_t = w.pad_t; a.pad_l = w.pad_l;
        a.TW = w.Wq > 128 ? 128 : w.Wq;
        switch (w.Cb) {
            case 16: return launch_thin<1, 16, 3>(a, st);
            case 32: return launch_thin<1, 32, 3>(a, st);
            case 48: return launch_thin<1, 48, 3>(a, st);
            default: return launch_thin<1, 64, 3>(a, st);
        }
    }
    const bool c28 = k3 && (w.Ca == 2 || w.Ca == 4) && w.Cb == 8;   // first layer of the 'pin' networks with one static variable (cfg5); cfg4's 4-channel tail
    if (!c28 && !((w.Ca == 1 || w.Ca == 8) && (w.Cb == 1 || w.Cb == 8))) return DL4DS_E_UNSUPPORTED;
    if (k7 && w.Ca == 8 && w.Cb == 8) return DL4DS_E_UNSUPPORTED;       // tensor-core kernels (conv_tc_wgrad2)
    if (w.Wq % 32 || (w.Wq > 128 && w.Wq % 128)) return DL4DS_E_UNSUPPORTED;
    if (w.Ca % 4 == 0 && (w.p_ld % 4 || (reinterpret_cast<uintptr_t>(w.P) & 15))) return DL4DS_E_UNSUPPORTED;
    if (w.Cb == 8 && (w.q_ld % 4 || (reinterpret_cast<uintptr_t>(w.Q) & 15))) return DL4DS_E_UNSUPPORTED;
    if (w.NQ < 16384) return DL4DS_E_UNSUPPORTED;          // tiny problems: the generic kernel is fine
    ThinWgradArgs a;
    a.P = w.P; a.Q = w.Q; a.dw = w.dw; a.p_ld = w.p_ld; a.q_ld = w.q_ld;
    a.N = w.N; a.H = w.Hq; a.W = w.Wq; a.pad_t = w.pad_t; a.pad_l = w.pad_l;
    a.TW = w.Wq > 128 ? 128 : w.Wq;
    if (k7) {
        if (w.Ca == 8 && w.Cb == 1) return launch_thin<8, 1, 7>(a, st);
        if (w.Ca == 1 && w.Cb == 8) return launch_thin<1, 8, 7>(a, st);
        return launch_thin<1, 1, 7>(a, st);
    }
    if (c28) return w.Ca == 2 ? launch_thin<2, 8, 3>(a, st) : launch_thin<4, 8, 3>(a, st);
    if (w.Ca == 8 && w.Cb == 8) return launch_thin<8, 8, 3>(a, st);
    if (w.Ca == 8 && w.Cb == 1) return launch_thin<8, 1, 3>(a, st);
    if (w.Ca == 1 && w.Cb == 8) return launch_thin<1, 8, 3>(a, st);
    return launch_thin<1, 1, 3>(a, st);
}

// -------------------------------------------------------------------------------------------------
// thin 3x3 direct convolution (forward, and input gradient = same op with flipped/transposed weights)
// for Cin, Cout in {1, 8}: one warp per output row, lane t owns pixels t, t+32, ... of the row (so every
// shared-memory access of a warp touches 32 consecutive pixels: conflict-free), all Cout outputs in
// registers; the weights sit in shared memory and are read as warp-uniform broadcasts.
// -------------------------------------------------------------------------------------------------
template <int CI, int CO, int KS>
__global__ void __launch_bounds__(256) thin_conv_kernel(ConvArgs p, int TW, int tiles_x, int tiles_y) {
    pdl_launch_dependents();    // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    p.x = pdl_after_wait(p.x);
    p.res = pdl_after_wait(p.res);
    constexpr int TH = 8, HALO = KS - 1, TAPS = KS * KS;
    extern __shared__ float sm[];
    __shared__ __align__(16) float ws[TAPS * CI * CO];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pitch = (TW + HALO) * CI + (CI == 8 ? 8 : (CI == 4 ? 4 : 1));   // row pitch (floats), rows land on distinct banks
    for (int i = tid; i < TAPS * CI * CO; i += 256) {
        const int co = i % CO, ci = (i / CO) % CI, tap = i / (CO * CI);
        ws[i] = (p.wmode == DL4DS_W_HWIO) ? __ldg(p.w + (tap * CI + ci) * CO + co)
                                           : __ldg(p.w + ((TAPS - 1 - tap) * CO + co) * CI + ci);
    }
    const int tile = blockIdx.x;
    const int img = tile / (tiles_x * tiles_y);
    const int trem = tile - img * tiles_x * tiles_y;
    const int ty = trem / tiles_x, tx = trem - ty * tiles_x;
    const int y0 = ty * TH, x0 = tx * TW;
    // ---- input tile with halo
    {
        const int cols = TW + HALO, rows = TH + HALO;
        if constexpr (CI % 4 == 0) {
            const int v4 = CI / 4, total = rows * cols * v4;
            for (int i = tid; i < total; i += 256) {
                const int c4 = i % v4, px = (i / v4) % cols, r = i / (v4 * cols);
                const int gy = y0 + r - p.pad_t, gx = x0 + px - p.pad_l;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W)
                    v = __ldg(reinterpret_cast<const float4*>(p.x + ((int64_t)(img * p.H + gy) * p.W + gx) * p.x_ld) + c4);
                *reinterpret_cast<float4*>(sm + (size_t)r * pitch + px * CI + c4 * 4) = v;
            }
        } else {
            const int total = rows * cols * CI;
            for (int i = tid; i < total; i += 256) {
                const int c = i % CI, px = (i / CI) % cols, r = i / (CI * cols);
                const int gy = y0 + r - p.pad_t, gx = x0 + px - p.pad_l;
                float v = 0.f;
                if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W)
                    v = __ldg(p.x + ((int64_t)(img * p.H + gy) * p.W + gx) * p.x_ld + c);
                sm[(size_t)r * pitch + px * CI + c] = v;
            }
        }
    }
    __syncthreads();
    const int oy = y0 + warp;
    if (oy >= p.H) return;
    const int npx = TW / 32;                                 // pixels per lane (1..4)
    // all of the lane's pixels at once: every weight read (a warp-uniform broadcast) feeds up to 4 pixels, which
    // cuts the shared-memory instructions per FMA by ~3x (r01b: 77 us for 8->8 at 128^2, LSU-bound)
    float acc[4][CO];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[j][c] = 0.f;
#pragma unroll
    for (int kh = 0; kh < KS; ++kh) {
        const float* row = sm + (size_t)(warp + kh) * pitch + lane * CI;
#pragma unroll
        for (int kw = 0; kw < KS; ++kw) {
            const float* wt = ws + (kh * KS + kw) * CI * CO;
            if constexpr (CI == 8) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {                 // 4 input channels at a time
                    float4 in[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (j < npx) in[j] = *reinterpret_cast<const float4*>(row + (32 * j + kw) * CI + 4 * h);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float wv[CO];
                        if constexpr (CO == 8) {
                            const float4 w0 = *reinterpret_cast<const float4*>(wt + (4 * h + u) * CO);
                            const float4 w1 = *reinterpret_cast<const float4*>(wt + (4 * h + u) * CO + 4);
                            wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w;
                            wv[4 % CO] = w1.x; wv[5 % CO] = w1.y; wv[6 % CO] = w1.z; wv[7 % CO] = w1.w;
                        } else {
#pragma unroll
                            for (int c = 0; c < CO; ++c) wv[c] = wt[(4 * h + u) * CO + c];
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (j < npx) {
                                const float x = u == 0 ? in[j].x : (u == 1 ? in[j].y : (u == 2 ? in[j].z : in[j].w));
                                if constexpr (CO == 8) {
                                    // packed fp32 FMA (FFMA2): 4 instructions + 1 pair setup for the 8 outputs
                                    const float2 xx = make_float2(x, x);
#pragma unroll
                                    for (int c = 0; c < CO; c += 2) {
                                        const float2 r = __ffma2_rn(xx, make_float2(wv[c], wv[c + 1 < CO ? c + 1 : c]),
                                                                    make_float2(acc[j][c], acc[j][c + 1 < CO ? c + 1 : c]));
                                        acc[j][c] = r.x;
                                        acc[j][c + 1 < CO ? c + 1 : c] = r.y;
                                    }
                                } else {
#pragma unroll
                                    for (int c = 0; c < CO; ++c) acc[j][c] = fmaf(x, wv[c], acc[j][c]);
                                }
                            }
                        }
                    }
                }
            } else {
#pragma unroll
                for (int ci = 0; ci < CI; ++ci) {
                    float wv[CO];
#pragma unroll
                    for (int c = 0; c < CO; ++c) wv[c] = wt[ci * CO + c];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j < npx) {
                            const float x = row[(32 * j + kw) * CI + ci];
#pragma unroll
                            for (int c = 0; c < CO; ++c) acc[j][c] = fmaf(x, wv[c], acc[j][c]);
                        }
                    }
                }
            }
        }
    }
    // ---- epilogue: bias, residual, activation, (accumulating) store
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j >= npx) break;
        const int xl = lane + 32 * j;
        const int64_t pix = ((int64_t)img * p.H + oy) * p.W + x0 + xl;
        float* yp = p.y + pix * p.y_ld;
        float o[CO];
#pragma unroll
        for (int c = 0; c < CO; ++c) {
            float v = acc[j][c];
            if (p.bias) v += __ldg(p.bias + c);
            if (p.res) v += __ldg(p.res + pix * p.res_ld + c);
            v = apply_act(v, p.act);
            if (p.beta) v += yp[c];
            o[c] = v;
        }
        if (CO == 8 && (p.y_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0)) {
            *reinterpret_cast<float4*>(yp) = make_float4(o[0], o[1 % CO], o[2 % CO], o[3 % CO]);
            *reinterpret_cast<float4*>(yp + 4) = make_float4(o[4 % CO], o[5 % CO], o[6 % CO], o[7 % CO]);
        } else {
#pragma unroll
            for (int c = 0; c < CO; ++c) yp[c] = o[c];
        }
    }
}

template <int CI, int CO, int KS>
static int launch_thin_conv(const ConvArgs& a, cudaStream_t st) {
    const int TW = a.W > 128 ? 128 : a.W;
    const int tiles_x = a.W / TW, tiles_y = (a.H + 7) / 8;
    const int pitch = (TW + KS - 1) * CI + (CI == 8 ? 8 : (CI == 4 ? 4 : 1));
    const size_t smem = (size_t)(8 + KS - 1) * pitch * 4;
    if (smem > 48 * 1024) {
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(thin_conv_kernel<CI, CO, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            attr = true;
        }
    }
    launch_pdl(8, thin_conv_kernel<CI, CO, KS>, dim3(a.N * tiles_x * tiles_y), dim3(256), smem, st, a, TW, tiles_x, tiles_y);
    return check_launch("thin_conv_kernel");
}

// -------------------------------------------------------------------------------------------------
// 3x3 convolution from ONE input channel to CO = 16 .. 64 channels (the first layer of the auxiliary / static-variable
// branch, ConvBlock_aux/conv1 in cfg3: 1 -> 48 at 128 x 128): a thread owns one pixel and all CO outputs; the nine
// inputs come from a shared halo tile, the weights are warp-uniform 16-byte broadcasts.  Write-bound (4*CO B/px).
// -------------------------------------------------------------------------------------------------
template <int CO>
__global__ void __launch_bounds__(256) thin_conv_c1_kernel(ConvArgs p, int tiles_x, int tiles_y) {
    pdl_launch_dependents();    // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    p.x = pdl_after_wait(p.x);
    constexpr int TH = 8, TW = 32;
    __shared__ float tile[TH + 2][TW + 4];
    __shared__ __align__(16) float ws[9 * CO];
    __shared__ __align__(16) float bs[CO];
    const int tid = threadIdx.x;
    for (int i = tid; i < 9 * CO; i += 256) {
        const int co = i % CO, tap = i / CO;
        ws[i] = (p.wmode == DL4DS_W_HWIO) ? __ldg(p.w + tap * CO + co) : __ldg(p.w + (8 - tap) * CO + co);
    }
    for (int i = tid; i < CO; i += 256) bs[i] = p.bias ? __ldg(p.bias + i) : 0.f;
    const int blk = blockIdx.x;
    const int img = blk / (tiles_x * tiles_y);
    const int trem = blk - img * tiles_x * tiles_y;
    const int ty = trem / tiles_x, tx = trem - ty * tiles_x;
    const int y0 = ty * TH, x0 = tx * TW;
    for (int i = tid; i < (TH + 2) * (TW + 2); i += 256) {
        const int r = i / (TW + 2), c = i - r * (TW + 2);
        const int gy = y0 + r - p.pad_t, gx = x0 + c - p.pad_l;
        float v = 0.f;
        if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) v = __ldg(p.x + ((int64_t)(img * p.H + gy) * p.W + gx) * p.x_ld);
        tile[r][c] = v;
    }
    __syncthreads();
    const int ry = tid >> 5, rx = tid & 31;
    const int oy = y0 + ry;
    if (oy >= p.H) return;
    float xin[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) xin[kh * 3 + kw] = tile[ry + kh][rx + kw];
    float* yp = p.y + (((int64_t)img * p.H + oy) * p.W + x0 + rx) * p.y_ld;
#pragma unroll
    for (int c4 = 0; c4 < CO / 4; ++c4) {
        float4 acc = *reinterpret_cast<const float4*>(bs + c4 * 4);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const float4 w = *reinterpret_cast<const float4*>(ws + tap * CO + c4 * 4);
            acc.x = fmaf(xin[tap], w.x, acc.x); acc.y = fmaf(xin[tap], w.y, acc.y);
            acc.z = fmaf(xin[tap], w.z, acc.z); acc.w = fmaf(xin[tap], w.w, acc.w);
        }
        acc.x = apply_act(acc.x, p.act); acc.y = apply_act(acc.y, p.act);
        acc.z = apply_act(acc.z, p.act); acc.w = apply_act(acc.w, p.act);
        *reinterpret_cast<float4*>(yp + c4 * 4) = acc;
    }
}

template <int CO>
static int launch_thin_c1(const ConvArgs& a, cudaStream_t st) {
    const int tiles_x = a.W / 32, tiles_y = (a.H + 7) / 8;
    launch_pdl(8, thin_conv_c1_kernel<CO>, dim3(a.N * tiles_x * tiles_y), dim3(256), 0, st, a, tiles_x, tiles_y);
    return check_launch("thin_conv_c1_kernel");
}

static int conv2d_fwd_thin_c1(const ConvArgs& a, cudaStream_t st) {
    if (a.KH != 3 || a.KW != 3 || a.stride != 1 || a.up != 1 || a.Ho != a.H || a.Wo != a.W || a.d2s_r > 1)
        return DL4DS_E_UNSUPPORTED;
    if (a.Cin != 1 || a.res || a.beta || a.W % 32) return DL4DS_E_UNSUPPORTED;
    if (a.y_ld % 4 || (reinterpret_cast<uintptr_t>(a.y) & 15)) return DL4DS_E_UNSUPPORTED;
    if ((int64_t)a.N * a.H * a.W < 16384) return DL4DS_E_UNSUPPORTED;
    switch (a.Cout) {
        case 16: return launch_thin_c1<16>(a, st);
        case 32: return launch_thin_c1<32>(a, st);
        case 48: return launch_thin_c1<48>(a, st);
        case 64: return launch_thin_c1<64>(a, st);
        default: return DL4DS_E_UNSUPPORTED;
    }
}

// DL4DS_E_UNSUPPORTED when the shape is outside this kernel's domain
int conv2d_fwd_thin(const ConvArgs& a, cudaStream_t st) {
    {
        const int rc = conv2d_fwd_thin_c1(a, st);       // 1 -> 16 / 32 / 48 / 64 channels
        if (rc != DL4DS_E_UNSUPPORTED) return rc;
    }
    const bool k3 = a.KH == 3 && a.KW == 3, k7 = a.KH == 7 && a.KW == 7;      // 7x7: the ConvNeXt stem / tail
    if (!(k3 || k7) || a.stride != 1 || a.up != 1 || a.Ho != a.H || a.Wo != a.W || a.d2s_r > 1)
        return DL4DS_E_UNSUPPORTED;
    const bool c28 = k3 && ((a.Cin == 2 && a.Cout == 8) || (a.Cin == 4 && a.Cout == 8) || (a.Cin == 8 && a.Cout == 4));      // first layer of the pin networks with one static variable (cfg5)
    if (!c28 && !((a.Cin == 1 || a.Cin == 8) && (a.Cout == 1 || a.Cout == 8))) return DL4DS_E_UNSUPPORTED;
    if (a.W % 32 || (a.W > 128 && a.W % 128)) return DL4DS_E_UNSUPPORTED;
    if (a.Cin % 4 == 0 && !a.vec) return DL4DS_E_UNSUPPORTED;
    if ((int64_t)a.N * a.H * a.W < 16384) return DL4DS_E_UNSUPPORTED;
    if (k7) {
        if (a.Cin == 8 && a.Cout == 8) return DL4DS_E_UNSUPPORTED;      // the tensor-core halo kernel
        if (a.Cin == 8 && a.Cout == 1) return launch_thin_conv<8, 1, 7>(a, st);
        if (a.Cin == 1 && a.Cout == 8) return launch_thin_conv<1, 8, 7>(a, st);
        return launch_thin_conv<1, 1, 7>(a, st);
    }
    if (c28) {
        if (a.Cin == 2) return launch_thin_conv<2, 8, 3>(a, st);
        if (a.Cin == 4) return launch_thin_conv<4, 8, 3>(a, st);
        return launch_thin_conv<8, 4, 3>(a, st);
    }
    if (a.Cin == 8 && a.Cout == 8) return launch_thin_conv<8, 8, 3>(a, st);
    if (a.Cin == 8 && a.Cout == 1) return launch_thin_conv<8, 1, 3>(a, st);
    if (a.Cin == 1 && a.Cout == 8) return launch_thin_conv<1, 8, 3>(a, st);
    return launch_thin_conv<1, 1, 3>(a, st);
}

// -------------------------------------------------------------------------------------------------
// pointwise (1x1) convolution for narrow layers (TransitionLast 48 -> 8 at 128 x 128 and its input
// gradient 8 -> 48, sp_postups.py:205; the 1x1 projections of the residual blocks): pure streaming.
// One thread = one 16-byte piece of the output (a pixel's 4 consecutive output channels): the warp's
// stores are one contiguous run, the pixel's input row is read straight from global memory (the
// threads of a pixel hit the same lines in L1), the weights sit in shared memory.  No barrier in the
// loop.  Measured (round 1): 69 us for 48 -> 8 and 103 us for 8 -> 48 over 2^20 pixels (a warp-tile variant
// staging 32 pixels per warp through shared memory was slower: 181 / 119 us).
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pointwise_conv_kernel(ConvArgs p, int n_items, int G) {
    pdl_launch_dependents();    // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    p.x = pdl_after_wait(p.x);
    p.res = pdl_after_wait(p.res);
    extern __shared__ __align__(16) float psm[];
    const int Cin = p.Cin, Cout = p.Cout;
    // channel counts that are no multiple of 4 (98 = 48 + 48 + 2 of cfg3's concatenation): the weight image is padded
    // with zeros to Cinp x Coutp, the last input group is masked, the last output group is stored lane by lane
    const int Coutp = G * 4, v4in = (Cin + 3) >> 2, Cinp = v4in * 4;
    const int in_tail = Cin & 3, out_tail = Cout & 3;
    float* wsm = psm;                                        // Cinp x Coutp
    float* bsm = wsm + Cinp * Coutp;                         // Coutp
    for (int i = threadIdx.x; i < Cinp * Coutp; i += 256) {
        const int co = i % Coutp, ci = i / Coutp;
        float w = 0.f;
        if (ci < Cin && co < Cout) w = (p.wmode == DL4DS_W_HWIO) ? __ldg(p.w + ci * Cout + co) : __ldg(p.w + co * Cin + ci);
        wsm[i] = w;
    }
    for (int i = threadIdx.x; i < Coutp; i += 256) bsm[i] = (p.bias && i < Cout) ? __ldg(p.bias + i) : 0.0f;
    __syncthreads();
    for (int idx = blockIdx.x * 256 + threadIdx.x; idx < n_items; idx += gridDim.x * 256) {
        const int px = idx / G, g = idx - px * G;
        const float* xrow = p.x + (int64_t)px * p.x_ld;
        const float4* xr = reinterpret_cast<const float4*>(xrow);
        const float* wg = wsm + g * 4;
        float4 acc = *reinterpret_cast<const float4*>(bsm + g * 4);
#pragma unroll 4
        for (int c4 = 0; c4 < v4in; ++c4) {
            float4 xv;
            if (in_tail && c4 == v4in - 1) {                 // (never reads past the row's Cin channels)
                xv.x = __ldg(xrow + c4 * 4);
                xv.y = in_tail > 1 ? __ldg(xrow + c4 * 4 + 1) : 0.f;
                xv.z = in_tail > 2 ? __ldg(xrow + c4 * 4 + 2) : 0.f;
                xv.w = 0.f;
            } else {
                xv = __ldg(xr + c4);
            }
            const float4 w0 = *reinterpret_cast<const float4*>(wg + (c4 * 4 + 0) * Coutp);
            const float4 w1 = *reinterpret_cast<const float4*>(wg + (c4 * 4 + 1) * Coutp);
            const float4 w2 = *reinterpret_cast<const float4*>(wg + (c4 * 4 + 2) * Coutp);
            const float4 w3 = *reinterpret_cast<const float4*>(wg + (c4 * 4 + 3) * Coutp);
            acc.x = fmaf(xv.x, w0.x, acc.x); acc.y = fmaf(xv.x, w0.y, acc.y); acc.z = fmaf(xv.x, w0.z, acc.z); acc.w = fmaf(xv.x, w0.w, acc.w);
            acc.x = fmaf(xv.y, w1.x, acc.x); acc.y = fmaf(xv.y, w1.y, acc.y); acc.z = fmaf(xv.y, w1.z, acc.z); acc.w = fmaf(xv.y, w1.w, acc.w);
            acc.x = fmaf(xv.z, w2.x, acc.x); acc.y = fmaf(xv.z, w2.y, acc.y); acc.z = fmaf(xv.z, w2.z, acc.z); acc.w = fmaf(xv.z, w2.w, acc.w);
            acc.x = fmaf(xv.w, w3.x, acc.x); acc.y = fmaf(xv.w, w3.y, acc.y); acc.z = fmaf(xv.w, w3.z, acc.z); acc.w = fmaf(xv.w, w3.w, acc.w);
        }
        if (out_tail && g == G - 1) {                        // partial last group: lane by lane
            float v[4] = {acc.x, acc.y, acc.z, acc.w};
            for (int l = 0; l < out_tail; ++l) {
                float o = v[l];
                if (p.res) o += __ldg(p.res + (int64_t)px * p.res_ld + g * 4 + l);
                o = apply_act(o, p.act);
                float* d = p.y + (int64_t)px * p.y_ld + g * 4 + l;
                if (p.beta) o += *d;
                *d = o;
            }
            continue;
        }
        if (p.res) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(p.res + (int64_t)px * p.res_ld) + g);
            acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
        }
        acc.x = apply_act(acc.x, p.act); acc.y = apply_act(acc.y, p.act);
        acc.z = apply_act(acc.z, p.act); acc.w = apply_act(acc.w, p.act);
        float4* dst = reinterpret_cast<float4*>(p.y + (int64_t)px * p.y_ld) + g;
        if (p.beta) {
            const float4 o = *dst;
            acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        }
        *dst = acc;
    }
}

static int launch_pointwise(const ConvArgs& a, cudaStream_t st) {
    const int64_t n_pix = (int64_t)a.N * a.H * a.W;
    const int G = (a.Cout + 3) / 4;
    const int64_t n_items = n_pix * G;
    if (n_items >= (1ll << 31)) return DL4DS_E_UNSUPPORTED;
    const size_t smem = (size_t)(((a.Cin + 3) / 4 * 4) * G * 4 + G * 4) * 4;
    if (smem > 48 * 1024) return DL4DS_E_UNSUPPORTED;
    int64_t blocks = (n_items + 255) / 256;
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    launch_pdl(8, pointwise_conv_kernel, dim3((unsigned)blocks), dim3(256), smem, st, a, (int)n_items, G);
    return check_launch("pointwise_conv_kernel");
}

// DL4DS_E_UNSUPPORTED when the shape is outside this kernel's domain
int conv2d_fwd_pointwise(const ConvArgs& a, cudaStream_t st) {
    if (a.KH != 1 || a.KW != 1 || a.stride != 1 || a.up != 1 || a.Ho != a.H || a.Wo != a.W || a.d2s_r > 1)
        return DL4DS_E_UNSUPPORTED;
    // rows must be 16-byte aligned (pitch % 4); the channel COUNTS may have a tail (round 2)
    // (a tensor with fewer than 4 channels has no full group: it is read / written lane by lane, no alignment needed)
    if (a.Cin > 128) return DL4DS_E_UNSUPPORTED;
    if (a.Cin >= 4 && (a.x_ld % 4 || (reinterpret_cast<uintptr_t>(a.x) & 15))) return DL4DS_E_UNSUPPORTED;
    if (a.Cin > 8 && a.Cout > 8) return DL4DS_E_UNSUPPORTED;          // wide x wide goes to the tensor cores
    if (a.Cout >= 4 && (a.y_ld % 4 || (reinterpret_cast<uintptr_t>(a.y) & 15))) return DL4DS_E_UNSUPPORTED;
    if (a.Cout >= 4 && a.res && (a.res_ld % 4 || (reinterpret_cast<uintptr_t>(a.res) & 15))) return DL4DS_E_UNSUPPORTED;
    if ((int64_t)a.N * a.H * a.W < 65536) return DL4DS_E_UNSUPPORTED;
    if (a.Cout > 128 || a.Cin < 2 || a.Cout < 2) return DL4DS_E_UNSUPPORTED;
    return launch_pointwise(a, st);
}

// -------------------------------------------------------------------------------------------------
// pointwise (1x1) weight gradient for narrow layers: dw[ca][cb] += sum_px P[px][ca] * Q[px][cb] with
// Ca*Cb/4 <= 256 (TransitionLast 48 x 8 over a million pixels, the first residual-block projections).
// A 128-pixel tile of P and Q is staged in shared memory with coalesced 16-byte loads; a thread owns one
// (ca, 4 cb) accumulator group and walks the pixels of its stream; several blocks per SM overlap the
// loads of one tile with the FMAs of another.  (The tensor-core kernels need >= 119 cycles per MMA whatever
// its size: r01b measured 150 us for this layer against a 36 us HBM roofline.)
// -------------------------------------------------------------------------------------------------
constexpr int kPwgTile = 128;

__global__ void __launch_bounds__(256) pointwise_wgrad_kernel(const float* __restrict__ P, int p_ld,
                                                              const float* __restrict__ Q, int q_ld,
                                                              float* __restrict__ dw, int64_t n_pix, int Ca, int Cb) {
    pdl_launch_dependents();    // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    P = pdl_after_wait(P);
    Q = pdl_after_wait(Q);
    extern __shared__ __align__(16) float wsm[];
    // channel counts with a tail (Ca = 98, Cb = 2 in cfg3): shared rows are padded to multiples of 4, the tail group is
    // loaded lane by lane (zeros beyond the tensor's channels), lanes beyond Cb are dropped at the end
    const int va = (Ca + 3) >> 2, G = (Cb + 3) >> 2, Cap = va * 4, Cbp = G * 4;
    const int a_tail = Ca & 3, b_tail = Cb & 3;
    float* ps = wsm;                                         // kPwgTile x Cap
    float* qs = ps + kPwgTile * Cap;                         // kPwgTile x Cbp
    float* red = qs + kPwgTile * Cbp;                        // Ca x Cbp
    const int tid = threadIdx.x;
    const int tps = Ca * G;                                  // threads per pixel stream
    const int streams = 256 / tps;
    const int stream = tid / tps, sub = tid - stream * tps;
    const int ca = sub / G, g = sub - ca * G;
    const bool active = stream < streams;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < Ca * Cbp; i += 256) red[i] = 0.0f;
    const int64_t ntiles = (n_pix + kPwgTile - 1) / kPwgTile;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * kPwgTile;
        const int npx = (int)min((int64_t)kPwgTile, n_pix - base);
        __syncthreads();
        for (int i = tid; i < npx * va; i += 256) {
            const int px = i / va, c4 = i - px * va;
            const float* src = P + (base + px) * p_ld + c4 * 4;
            float4 v;
            if (a_tail && c4 == va - 1) {
                v.x = __ldg(src); v.y = a_tail > 1 ? __ldg(src + 1) : 0.f; v.z = a_tail > 2 ? __ldg(src + 2) : 0.f; v.w = 0.f;
            } else {
                v = __ldg(reinterpret_cast<const float4*>(src));
            }
            *reinterpret_cast<float4*>(ps + px * Cap + c4 * 4) = v;
        }
        for (int i = tid; i < npx * G; i += 256) {
            const int px = i / G, c4 = i - px * G;
            const float* src = Q + (base + px) * q_ld + c4 * 4;
            float4 v;
            if (b_tail && c4 == G - 1) {
                v.x = __ldg(src); v.y = b_tail > 1 ? __ldg(src + 1) : 0.f; v.z = b_tail > 2 ? __ldg(src + 2) : 0.f; v.w = 0.f;
            } else {
                v = __ldg(reinterpret_cast<const float4*>(src));
            }
            *reinterpret_cast<float4*>(qs + px * Cbp + c4 * 4) = v;
        }
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (int px = stream; px < npx; px += streams) {
                const float x = ps[px * Cap + ca];
                const float4 q = *reinterpret_cast<const float4*>(qs + px * Cbp + g * 4);
                acc.x = fmaf(x, q.x, acc.x); acc.y = fmaf(x, q.y, acc.y);
                acc.z = fmaf(x, q.z, acc.z); acc.w = fmaf(x, q.w, acc.w);
            }
        }
    }
    __syncthreads();
    if (active) {
        float* r = red + ca * Cbp + g * 4;
        atomicAdd(r + 0, acc.x); atomicAdd(r + 1, acc.y); atomicAdd(r + 2, acc.z); atomicAdd(r + 3, acc.w);
    }
    __syncthreads();
    for (int i = tid; i < Ca * Cbp; i += 256) {
        const int a_ = i / Cbp, b_ = i - a_ * Cbp;
        if (b_ < Cb) atomicAdd(dw + a_ * Cb + b_, red[i]);
    }
}

// DL4DS_E_UNSUPPORTED when the shape is outside this kernel's domain
int conv2d_wgrad_pointwise(const WgradArgs& w, cudaStream_t st) {
    if (w.KH != 1 || w.KW != 1 || w.stride != 1 || w.Hp != w.Hq || w.Wp != w.Wq) return DL4DS_E_UNSUPPORTED;
    if (w.Ca < 2 || w.Ca * ((w.Cb + 3) / 4) > 256) return DL4DS_E_UNSUPPORTED;
    if (w.Ca >= 4 && (w.p_ld % 4 || (reinterpret_cast<uintptr_t>(w.P) & 15))) return DL4DS_E_UNSUPPORTED;
    if (w.Cb >= 4 && (w.q_ld % 4 || (reinterpret_cast<uintptr_t>(w.Q) & 15))) return DL4DS_E_UNSUPPORTED;
    if (w.NQ < 16384) return DL4DS_E_UNSUPPORTED;
    const int Cap = (w.Ca + 3) / 4 * 4, Cbp = (w.Cb + 3) / 4 * 4;
    const size_t smem = (size_t)(kPwgTile * (Cap + Cbp) + w.Ca * Cbp) * 4;
    if (smem > 96 * 1024) return DL4DS_E_UNSUPPORTED;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(pointwise_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        attr = true;
    }
    const int64_t ntiles = (w.NQ + kPwgTile - 1) / kPwgTile;
    int blocks_per_sm = (int)((200 * 1024) / (smem + 1024));   // measured: 6 per SM beats 3 (94 vs 115 us)
    if (blocks_per_sm > 8) blocks_per_sm = 8;
    int64_t grid = (int64_t)kNumSMs * blocks_per_sm;
    if (grid > ntiles) grid = ntiles;
    launch_pdl(8, pointwise_wgrad_kernel, dim3((unsigned)grid), dim3(256), smem, st, w.P, w.p_ld, w.Q, w.q_ld, w.dw, w.NQ, w.Ca, w.Cb);
    return check_launch("pointwise_wgrad_kernel");
}

// -------------------------------------------------------------------------------------------------
// vectorised bias / activation backward (+ space_to_depth un-shuffle)
// -------------------------------------------------------------------------------------------------
// block = (TX, PY): thread (tx, ty) owns float4 channel groups g = tx + k*TX (k < KS), pixels ty, ty+PY, ...
template <int KS>
__global__ void __launch_bounds__(256) bias_act_bwd_vec4_kernel(
    const float* __restrict__ dy, int dy_ld, const float* __restrict__ y, int y_ld,
    float* __restrict__ dz, int dz_ld, float* __restrict__ dbias,
    int64_t n_pix, int Ho, int Wo, int C, int act, int r) {
    pdl_launch_dependents();    // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    dy = pdl_after_wait(dy);
    y = pdl_after_wait(y);
    __shared__ float4 red[256];
    const int TX = blockDim.x, PY = blockDim.y;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int G = C / 4;
    float4 acc[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int Cd = (r > 1) ? C / (r * r) : C;
    if (KS == 1 && r > 1 && tx < G) {
        // depth_to_space un-shuffle (SubpixelConvolutionBlock, linear): pure permuting copy + column sums, 4 pixels
        // (independent 16-byte loads) in flight per thread; 32-bit index math (n_pix < 2^31)
        const int c = tx * 4;
        const int grp = c / Cd, cc = c - grp * Cd;
        const int di = grp / r, dj = grp - di * r;
        const int hw = Ho * Wo, Wr = Wo * r;
        const int step = gridDim.x * PY;
        const int npx = (int)n_pix;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p0 = blockIdx.x * PY + ty; p0 < npx; p0 += 4 * step) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int p = p0 + u * step;
                if (p < npx) {
                    const int n = p / hw, rem = p - n * hw;
                    const int oy = rem / Wo, ox = rem - oy * Wo;
                    const int64_t hp = ((int64_t)(n * Ho + oy) * r + di) * Wr + ox * r + dj;
                    v[u] = __ldg(reinterpret_cast<const float4*>(dy + hp * dy_ld + cc));
                    if (act != DL4DS_ACT_NONE) {
                        const float4 o = __ldg(reinterpret_cast<const float4*>(y + hp * y_ld + cc));
                        v[u].x *= act_grad_from_out(o.x, act); v[u].y *= act_grad_from_out(o.y, act);
                        v[u].z *= act_grad_from_out(o.z, act); v[u].w *= act_grad_from_out(o.w, act);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int p = p0 + u * step;
                if (p < npx) {
                    if (dz) *reinterpret_cast<float4*>(dz + (int64_t)p * dz_ld + c) = v[u];
                    a0.x += v[u].x; a0.y += v[u].y; a0.z += v[u].z; a0.w += v[u].w;
                }
            }
        }
        acc[0] = a0;
    } else if (KS == 1 && r == 1 && tx < G) {
        // same-layout case (every backbone layer): 4 pixels in flight per thread, loads of dy and y issued together
        const int c = tx * 4;
        const int64_t step = (int64_t)gridDim.x * PY;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int64_t p0 = (int64_t)blockIdx.x * PY + ty; p0 < n_pix; p0 += 4 * step) {
            float4 v[4], o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t p = p0 + u * step;
                if (p < n_pix) {
                    v[u] = __ldg(reinterpret_cast<const float4*>(dy + p * dy_ld + c));
                    if (act != DL4DS_ACT_NONE) o[u] = __ldg(reinterpret_cast<const float4*>(y + p * y_ld + c));
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t p = p0 + u * step;
                if (p < n_pix) {
                    if (act != DL4DS_ACT_NONE) {
                        v[u].x *= act_grad_from_out(o[u].x, act); v[u].y *= act_grad_from_out(o[u].y, act);
                        v[u].z *= act_grad_from_out(o[u].z, act); v[u].w *= act_grad_from_out(o[u].w, act);
                    }
                    if (dz) *reinterpret_cast<float4*>(dz + p * dz_ld + c) = v[u];
                    a0.x += v[u].x; a0.y += v[u].y; a0.z += v[u].z; a0.w += v[u].w;
                }
            }
        }
        acc[0] = a0;
    } else
    for (int64_t p = (int64_t)blockIdx.x * PY + ty; p < n_pix; p += (int64_t)gridDim.x * PY) {
        int n = 0, oy = 0, ox = 0;
        if (r > 1) {
            const int hw = Ho * Wo;
            n = (int)(p / hw);
            const int rem = (int)(p - (int64_t)n * hw);
            oy = rem / Wo; ox = rem - oy * Wo;
        }
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            const int g = tx + k * TX;
            if (g < G) {
                const int c = g * 4;
                float4 v;
                if (r > 1) {
                    const int grp = c / Cd, cc = c - grp * Cd;
                    const int di = grp / r, dj = grp - di * r;
                    const int64_t hp = ((int64_t)(n * Ho * r + oy * r + di)) * (Wo * r) + ox * r + dj;
                    v = __ldg(reinterpret_cast<const float4*>(dy + hp * dy_ld + cc));
                    if (act != DL4DS_ACT_NONE) {
                        const float4 o = __ldg(reinterpret_cast<const float4*>(y + hp * y_ld + cc));
                        v.x *= act_grad_from_out(o.x, act); v.y *= act_grad_from_out(o.y, act);
                        v.z *= act_grad_from_out(o.z, act); v.w *= act_grad_from_out(o.w, act);
                    }
                } else {
                    v = __ldg(reinterpret_cast<const float4*>(dy + p * dy_ld + c));
                    if (act != DL4DS_ACT_NONE) {
                        const float4 o = __ldg(reinterpret_cast<const float4*>(y + p * y_ld + c));
                        v.x *= act_grad_from_out(o.x, act); v.y *= act_grad_from_out(o.y, act);
                        v.z *= act_grad_from_out(o.z, act); v.w *= act_grad_from_out(o.w, act);
                    }
                }
                if (dz) *reinterpret_cast<float4*>(dz + p * dz_ld + c) = v;
                acc[k].x += v.x; acc[k].y += v.y; acc[k].z += v.z; acc[k].w += v.w;
            }
        }
    }
    if (dbias == nullptr) return;
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        red[ty * TX + tx] = acc[k];
        __syncthreads();
        if (ty == 0) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int j = 0; j < PY; ++j) {
                const float4 t = red[j * TX + tx];
                s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
            }
            const int g = tx + k * TX;
            if (g < G) {
                atomicAdd(dbias + g * 4 + 0, s.x); atomicAdd(dbias + g * 4 + 1, s.y);
                atomicAdd(dbias + g * 4 + 2, s.z); atomicAdd(dbias + g * 4 + 3, s.w);
            }
        }
        __syncthreads();
    }
}

int bias_act_bwd_vec4(const float* dy, int dy_ld, const float* y, int y_ld, float* dz, int dz_ld, float* dbias,
                      int64_t n_pix, int Ho, int Wo, int C, int act, int r, cudaStream_t st) {
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const int Cd = r > 1 ? C / (r * r) : C;
    if (C % 4 || Cd % 4 || dy_ld % 4 || !al(dy)) return DL4DS_E_UNSUPPORTED;
    if (dz && (dz_ld % 4 || !al(dz))) return DL4DS_E_UNSUPPORTED;
    if (act != DL4DS_ACT_NONE && (y_ld % 4 || !al(y))) return DL4DS_E_UNSUPPORTED;
    const int G = C / 4;
    int TX = 1;
    while (TX < G && TX < 64) TX <<= 1;
    if (G <= 64 && (G & (G - 1)) != 0) TX = G;              // e.g. 48, 12, 10 or 6 float4 groups: no idle lanes
    const int KS = (G + TX - 1) / TX;
    if (KS > 4) return DL4DS_E_UNSUPPORTED;
    const int PY = 256 / TX;
    dim3 block(TX, PY);
    int64_t want = (n_pix + PY * 4 - 1) / (PY * 4);
    int grid = (int)(want > 8 * kNumSMs ? 8 * kNumSMs : (want < 1 ? 1 : want));
#define LAUNCH_V4(K) launch_pdl(8, bias_act_bwd_vec4_kernel<K>, dim3(grid), dim3(block), 0, st, dy, dy_ld, y, y_ld, dz, dz_ld, dbias, n_pix, Ho, Wo, C, act, r)
    if (KS == 1) LAUNCH_V4(1);
    else if (KS == 2) LAUNCH_V4(2);
    else LAUNCH_V4(4);
#undef LAUNCH_V4
    return check_launch("bias_act_bwd_vec4");
}

}  // namespace dl4ds
