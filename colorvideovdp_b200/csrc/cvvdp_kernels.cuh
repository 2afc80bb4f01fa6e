// cvvdp_kernels.cuh -- sm_100a kernels of the ColorVideoVDP hot path.
//
// HBM layout of every pyramid level: planes of float4 "pixels", one plane per (batch item, frame,
// video) with video 0 = test, 1 = reference; a pixel holds the four perceptual channels
// (A-sust, RG, YV, A-trans) of that video, so each pixel is one aligned 128-bit load/store:
//     level_i[((b * n + f) * 2 + v) * h_i * w_i + y * w_i + x]          (float4)
// (the reference keeps [B, 8, N, H, W] planar fp32 with test/ref interleaved on the channel axis,
// cvvdp_metric.py:550-560).  References in comments are paths in the reference tree.
#pragma once
#include "cvvdp_common.cuh"

namespace cvvdp {

// =================================================================================================
// Front end: dtype unpack -> EOTF -> RGB->DKLd65   (video_source.py:320-346, display_model.py:333-365,
// 241-276)
// =================================================================================================
__device__ __forceinline__ float load_unpack(const void *base, long long off, int dtype) {
    switch (dtype) {
        case CVVDP_DTYPE_U8: return (float)((const unsigned char *)base)[off] / 255.0f;
        case CVVDP_DTYPE_U16: return (float)((const unsigned short *)base)[off] * (1.0f / 65535.0f);
        case CVVDP_DTYPE_F16: return half_bits_to_float(((const unsigned short *)base)[off]);
        default: return ((const float *)base)[off];
    }
}

__device__ __forceinline__ float clamp01_keepnan(float v) { return v < 0.f ? 0.f : (v > 1.f ? 1.f : v); }

__device__ __forceinline__ float srgb2lin(float p) {  // display_model.py:78-80
    return p > 0.04045f ? f_pow((p + 0.055f) * (1.0f / 1.055f), 2.4f) : p * (1.0f / 12.92f);
}
__device__ __forceinline__ float pq2lin(float V) {  // display_model.py:58-70
    const float c1 = 0.8359375f, c2 = 18.8515625f, c3 = 18.6875f;
    float im_t = f_pow(V, (float)(1.0 / 78.84375));
    float num = fmaxf(im_t - c1, 0.f);
    return 10000.0f * f_pow(num / (c2 - c3 * im_t), (float)(1.0 / 0.1593017578125));
}

// v[0..N) display-encoded -> absolute linear (cd/m^2), in place.  N = 1 or 3 is a compile-time count: with a
// run-time one the small array is indexed dynamically and lands in local memory (measured in round 2: the fp32
// input path of the temporal kernel ran 8x slower than the 8-bit table path because of it).
template <int N>
__device__ __forceinline__ void eotf_forward_n(float (&v)[3], const DisplayDev &d) {
    const float a = d.Ypeak - d.Yblack;
    if (d.eotf == CVVDP_EOTF_NONE) return;
    if (d.eotf != CVVDP_EOTF_LINEAR) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = clamp01_keepnan(v[i]);  // display_model.py:335-337
    }
    switch (d.eotf) {
        case CVVDP_EOTF_SRGB:
#pragma unroll
            for (int i = 0; i < N; ++i) {
                float lin = srgb2lin(v[i]);
                if (d.exposure != 1.f) lin = fminf(fmaxf(lin * d.exposure, 0.f), 1.f);
                v[i] = a * lin + d.Yblack + d.Yrefl;
            }
            break;
        case CVVDP_EOTF_PQ:
#pragma unroll
            for (int i = 0; i < N; ++i)
                v[i] = fminf(fmaxf(pq2lin(v[i]) * d.exposure, 0.005f), d.Ypeak) + d.Yblack + d.Yrefl;
            break;
        case CVVDP_EOTF_LINEAR:
#pragma unroll
            for (int i = 0; i < N; ++i) v[i] = fminf(fmaxf(v[i] * d.exposure, d.lin_lo), d.Ypeak) + d.Yrefl;
            break;
        case CVVDP_EOTF_HLG: {  // display_model.py:89-108, 350-359 (needs all three channels; the plan refuses N = 1)
            // exp / pow through the MUFU pair like every other power on this path: the libm versions cost ~100
            // instructions per call and, unrolled over pixels and frames, blew the float-input kernel up to 240 KB
            const float ha = 0.17883277f, hb = 1.f - 4.f * ha, hc = 0.5f - ha * -0.33500979f;  // ln(4 ha) = -0.33500979
            const float k_exp = 1.4426950408889634f / ha;
            float s[3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
                s[i] = v[i] <= 0.5f ? v[i] * v[i] / 3.0f : (f_ex2((v[i] - hc) * k_exp) + hb) / 12.0f;
            float Ys = 0.2627f * s[0] + 0.6780f * s[1] + 0.0593f * s[2];
            float gsc = f_pow(Ys, d.gamma - 1.f);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                float lin = gsc * s[i];
                if (d.exposure != 1.f) lin = fminf(fmaxf(lin * d.exposure, 0.f), 1.f);
                v[i] = a * lin + d.Yblack + d.Yrefl;
            }
        } break;
        default:  // CVVDP_EOTF_GAMMA, display_model.py:360-362
#pragma unroll
            for (int i = 0; i < N; ++i) {
                float lin = fminf(fmaxf(f_pow(v[i], d.gamma) * d.exposure, 0.f), 1.f);
                v[i] = a * lin + d.Yblack + d.Yrefl;
            }
    }
}
__device__ __forceinline__ void eotf_forward(float (&v)[3], int n, const DisplayDev &d) {
    if (n == 3) eotf_forward_n<3>(v, d);
    else eotf_forward_n<1>(v, d);
}
// The same arithmetic for an EOTF known at compile time (sRGB, PQ or linear), three values: the temporal kernel picks
// the body once per chunk of frames instead of branching per value (the run-time switch, taken twice per pixel pair
// and frame, and the dtype switch next to it were a third of the float-input kernel's instructions).
template <int E, bool IN_RANGE = false>  // IN_RANGE: the values are known to lie in [0, 1] (decoded YUV): no clamp needed
__device__ __forceinline__ void eotf_forward_c(float (&v)[3], const DisplayDev &d) {
    const float a = d.Ypeak - d.Yblack;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (E == CVVDP_EOTF_SRGB) {
            float lin = srgb2lin(IN_RANGE ? v[i] : clamp01_keepnan(v[i]));
            if (d.exposure != 1.f) lin = fminf(fmaxf(lin * d.exposure, 0.f), 1.f);
            v[i] = a * lin + d.Yblack + d.Yrefl;
        } else if (E == CVVDP_EOTF_PQ) {
            v[i] = fminf(fmaxf(pq2lin(IN_RANGE ? v[i] : clamp01_keepnan(v[i])) * d.exposure, 0.005f), d.Ypeak) + d.Yblack + d.Yrefl;
        } else {  // CVVDP_EOTF_LINEAR
            v[i] = fminf(fmaxf(v[i] * d.exposure, d.lin_lo), d.Ypeak) + d.Yrefl;
        }
    }
}

// Planar YUV pixel -> display-encoded RGB in 0..1 (video_source_yuv.py:153-233): limited-range unpack,
// bilinear chroma upsampling with torch's align_corners=False rule (src = (dst + .5)/2 - .5, clamped at
// 0, neighbour index clamped to the last sample), YCbCr->RGB, clip.  fbase: element offset of the frame.
__device__ __forceinline__ float yuv_sample(const void *data, long long off, int dtype) {
    return dtype == CVVDP_DTYPE_U8 ? (float)((const unsigned char *)data)[off] : (float)((const unsigned short *)data)[off];
}
// chroma samples (columns i0, i1, rows j0, j1) and weights of luma pixel (y, x)
__device__ __forceinline__ void yuv_chroma_taps(const YuvDev &yu, int y, int x, int &i0, int &i1, int &j0, int &j1, float &lx,
                                                float &ly) {
    const int cw = yu.chroma == 444 ? yu.W : yu.W / 2, ch = yu.chroma == 420 ? yu.H / 2 : yu.H;
    i0 = i1 = x;
    j0 = j1 = y;
    lx = ly = 0.f;
    if (yu.chroma != 444) {
        const float sx = fmaxf(((float)x + 0.5f) * 0.5f - 0.5f, 0.f);
        i0 = (int)sx;
        lx = sx - (float)i0;
        i1 = min(i0 + 1, cw - 1);
    }
    if (yu.chroma == 420) {
        const float sy = fmaxf(((float)y + 0.5f) * 0.5f - 0.5f, 0.f);
        j0 = (int)sy;
        ly = sy - (float)j0;
        j1 = min(j0 + 1, ch - 1);
    }
}
__device__ __forceinline__ void yuv_fetch_rgb(const ClipView &cv, const YuvDev &yu, int dtype, long long fbase, int y, int x,
                                              float rgb[3]) {
    const long long ypix = (long long)yu.W * yu.H;
    const int cw = yu.chroma == 444 ? yu.W : yu.W / 2, ch = yu.chroma == 420 ? yu.H / 2 : yu.H;
    const long long uvpix = (long long)cw * ch;
    const float Y = fminf(fmaxf(yu.yw * yuv_sample(cv.data, fbase + (long long)y * yu.W + x, dtype) - yu.yo, 0.f), 1.f);
    int i0, i1, j0, j1;
    float lx, ly;
    yuv_chroma_taps(yu, y, x, i0, i1, j0, j1, lx, ly);
    float uv[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const long long pb = fbase + ypix + p * uvpix;
        const float c00 = fminf(fmaxf(yu.cw * yuv_sample(cv.data, pb + (long long)j0 * cw + i0, dtype) - yu.co, -0.5f), 0.5f);
        const float c01 = fminf(fmaxf(yu.cw * yuv_sample(cv.data, pb + (long long)j0 * cw + i1, dtype) - yu.co, -0.5f), 0.5f);
        const float c10 = fminf(fmaxf(yu.cw * yuv_sample(cv.data, pb + (long long)j1 * cw + i0, dtype) - yu.co, -0.5f), 0.5f);
        const float c11 = fminf(fmaxf(yu.cw * yuv_sample(cv.data, pb + (long long)j1 * cw + i1, dtype) - yu.co, -0.5f), 0.5f);
        uv[p] = (1.f - ly) * ((1.f - lx) * c00 + lx * c01) + ly * ((1.f - lx) * c10 + lx * c11);
    }
    rgb[0] = fminf(fmaxf(Y + yu.m_rv * uv[1], 0.f), 1.f);
    rgb[1] = fminf(fmaxf(Y + yu.m_gu * uv[0] + yu.m_gv * uv[1], 0.f), 1.f);
    rgb[2] = fminf(fmaxf(Y + yu.m_bu * uv[0], 0.f), 1.f);
}

// Input validation of the fused path (display_model.py:335-337, video_source.py:48-59): bit 0 = a value outside
// 0..1 reaches an EOTF that clamps, bit 1 = NaN, bit 2 = Inf.  Only floating-point clips can carry any of them.
__device__ __forceinline__ unsigned input_bits(float v, bool check_range) {
    unsigned f = (check_range && (v > 1.f || v < 0.f)) ? 1u : 0u;
    f |= (v != v) ? 2u : 0u;
    f |= (fabsf(v) == INFINITY) ? 4u : 0u;
    return f;
}
// Warp-aggregated publication of the validity bits: one atomic per set bit and warp, only when something is wrong.
__device__ __forceinline__ void publish_input_bits(unsigned bits, int *flags) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bits |= __shfl_xor_sync(0xffffffffu, bits, o);
    if ((threadIdx.x & 31) == 0 && bits != 0u && flags != nullptr) {
        if (bits & 1u) atomicAdd(&flags[0], 1);
        if (bits & 2u) atomicAdd(&flags[1], 1);
        if (bits & 4u) atomicAdd(&flags[2], 1);
    }
}

// Raw pixel of frame `fidx` (index inside the view) -> DKL triple.  vbits (optional) collects input_bits.
__device__ __forceinline__ void pixel_to_dkl(const ClipView &cv, long long base, int fidx, int cin, int dtype,
                                             const DisplayDev &d, float &o0, float &o1, float &o2,
                                             const YuvDev *yu = nullptr, int b = 0, int y = 0, int x = 0,
                                             unsigned *vbits = nullptr) {
    float v[3];
    const long long off = base + (long long)fidx * cv.s[2];
    if (yu != nullptr && yu->chroma != 0) {
        yuv_fetch_rgb(cv, *yu, dtype, b * cv.s[0] + (long long)fidx * cv.s[2], y, x, v);
    } else {
        v[0] = load_unpack(cv.data, off, dtype);
        if (cin == 3) {
            v[1] = load_unpack(cv.data, off + cv.s[1], dtype);
            v[2] = load_unpack(cv.data, off + 2 * cv.s[1], dtype);
        }
        if (vbits != nullptr && dtype >= CVVDP_DTYPE_F16) {
            const bool rng = d.eotf != CVVDP_EOTF_LINEAR && d.eotf != CVVDP_EOTF_NONE;
            for (int i = 0; i < cin; ++i) *vbits |= input_bits(v[i], rng);
        }
    }
    eotf_forward(v, cin, d);  // CVVDP_EOTF_NONE: values pass through (the host then sets M = identity)
    if (cin == 3) {  // display_model.py:266-269
        o0 = (v[0] * d.M[0] + v[1] * d.M[1]) + v[2] * d.M[2];
        o1 = (v[0] * d.M[3] + v[1] * d.M[4]) + v[2] * d.M[5];
        o2 = (v[0] * d.M[6] + v[1] * d.M[7]) + v[2] * d.M[8];
    } else {  // display_model.py:231-235 + the broadcast at cvvdp_metric.py:464-465, 503-504
        o0 = o1 = o2 = v[0];
    }
}

// cvvdp_metric.py:445-450
__device__ __forceinline__ int symmetric_frame_index(int fi, int F) {
    const int m = F - 1;
    const int a = -fi - 1;  // fi < 0
    if (((a / m) & 1) == 0) return (a % m) + 1;
    int r = fi % m;
    return r < 0 ? r + m : r;
}

// Standalone front end for the display-model plugin surface
// (vvdp_display_photometry.source_2_target_colorspace(frame, 'DKLd65')).
struct FrontendArgs {
    ClipView clip;
    DisplayDev dd;
    int dtype, cin, B, H, W, frame;
    YuvDev yuv;
    float *dst;     // [B, cin, H, W]
    int *flags;     // [0]: values outside 0..1, [1]: NaN, [2]: Inf  (may be null)
};
__global__ void __launch_bounds__(256) k_frontend(const FrontendArgs a) {
    const long long npix = (long long)a.H * a.W;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (p >= npix) return;
    const int y = (int)(p / a.W), x = (int)(p - (long long)y * a.W);
    const long long base = b * a.clip.s[0] + y * a.clip.s[3] + x * a.clip.s[4] + (long long)a.frame * a.clip.s[2];
    if (a.flags && a.yuv.chroma == 0) {
        for (int c = 0; c < a.cin; ++c) {
            float v = load_unpack(a.clip.data, base + c * a.clip.s[1], a.dtype);
            if (a.dd.eotf != CVVDP_EOTF_LINEAR && (v > 1.f || v < 0.f)) atomicAdd(&a.flags[0], 1);
            if (v != v) atomicAdd(&a.flags[1], 1);
            if (fabsf(v) == INFINITY) atomicAdd(&a.flags[2], 1);
        }
    }
    float o0, o1, o2;
    pixel_to_dkl(a.clip, base - (long long)a.frame * a.clip.s[2], a.frame, a.cin, a.dtype, a.dd, o0, o1, o2, &a.yuv, b, y, x);
    float *dst = a.dst + (long long)b * a.cin * npix + p;
    dst[0] = o0;
    if (a.cin == 3) {
        dst[npix] = o1;
        dst[2 * npix] = o2;
    }
}

// =================================================================================================
// Full-screen resize of a decoded frame (run_cvvdp.py:100 `--full-screen-resize`; video_source_yuv.py:257-260,333-336,
// video_source_file.py:280-285): torch.nn.functional.interpolate(size=..., mode=...) without align_corners or
// antialiasing, followed by clip(0, 1).  The interpolation rules are torch's (third-party to the reference):
//   nearest   src = min(floor(dst * in/out), in-1)
//   bilinear  src = max((dst + 0.5) * in/out - 0.5, 0); the second tap stays on the last sample at the edge
//   bicubic   src = (dst + 0.5) * in/out - 0.5, Keys kernel with A = -0.75, tap indices clamped to the plane
//   area      adaptive average pool: mean over [floor(i in/out), ceil((i+1) in/out))
// One thread per output pixel, all channels; the source planes are small enough to be served from L2.
// =================================================================================================
struct ResizeArgs {
    const float *src;  // [C][H][W]
    float *dst;        // [C][OH][OW]
    int C, H, W, OH, OW, mode, clip01;
};
__device__ __forceinline__ float resize_cubic1(float x) { return ((1.25f * x - 2.25f) * x) * x + 1.f; }            // |x| <= 1, A=-0.75
__device__ __forceinline__ float resize_cubic2(float x) { return ((-0.75f * x + 3.75f) * x - 6.f) * x + 3.f; }     // 1 < |x| < 2
__device__ __forceinline__ void resize_cubic_taps(float t, float (&c)[4]) {
    c[0] = resize_cubic2(t + 1.f);
    c[1] = resize_cubic1(t);
    c[2] = resize_cubic1(1.f - t);
    c[3] = resize_cubic2((1.f - t) + 1.f);
}
__global__ void __launch_bounds__(256) k_resize(const ResizeArgs a) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long onp = (long long)a.OH * a.OW, inp = (long long)a.H * a.W;
    if (p >= onp) return;
    const int oy = (int)(p / a.OW), ox = (int)(p - (long long)oy * a.OW);
    const float sy = (float)a.H / (float)a.OH, sx = (float)a.W / (float)a.OW;
    for (int c = 0; c < a.C; ++c) {
        const float *s = a.src + c * inp;
        float v;
        if (a.mode == CVVDP_RESIZE_NEAREST) {
            const int y = min((int)floorf(oy * sy), a.H - 1), x = min((int)floorf(ox * sx), a.W - 1);
            v = s[(long long)y * a.W + x];
        } else if (a.mode == CVVDP_RESIZE_BILINEAR) {
            const float ry = fmaxf(sy * (oy + 0.5f) - 0.5f, 0.f), rx = fmaxf(sx * (ox + 0.5f) - 0.5f, 0.f);
            const int y0 = min((int)ry, a.H - 1), x0 = min((int)rx, a.W - 1);
            const int y1 = y0 + (y0 < a.H - 1 ? 1 : 0), x1 = x0 + (x0 < a.W - 1 ? 1 : 0);
            const float ly = ry - y0, lx = rx - x0;
            const float top = (1.f - lx) * s[(long long)y0 * a.W + x0] + lx * s[(long long)y0 * a.W + x1];
            const float bot = (1.f - lx) * s[(long long)y1 * a.W + x0] + lx * s[(long long)y1 * a.W + x1];
            v = (1.f - ly) * top + ly * bot;
        } else if (a.mode == CVVDP_RESIZE_BICUBIC) {
            const float ry = sy * (oy + 0.5f) - 0.5f, rx = sx * (ox + 0.5f) - 0.5f;
            const float fy = floorf(ry), fx = floorf(rx);
            float cy[4], cx[4];
            resize_cubic_taps(ry - fy, cy);
            resize_cubic_taps(rx - fx, cx);
            v = 0.f;
            for (int i = 0; i < 4; ++i) {
                const int y = max(min((int)fy - 1 + i, a.H - 1), 0);
                float r = 0.f;
                for (int j = 0; j < 4; ++j) r += s[(long long)y * a.W + max(min((int)fx - 1 + j, a.W - 1), 0)] * cx[j];
                v += r * cy[i];
            }
        } else {  // area
            const int y0 = (int)(((long long)oy * a.H) / a.OH), y1 = (int)((((long long)oy + 1) * a.H + a.OH - 1) / a.OH);
            const int x0 = (int)(((long long)ox * a.W) / a.OW), x1 = (int)((((long long)ox + 1) * a.W + a.OW - 1) / a.OW);
            float sum = 0.f;
            for (int y = y0; y < y1; ++y)
                for (int x = x0; x < x1; ++x) sum += s[(long long)y * a.W + x];
            v = sum / (float)((y1 - y0) * (x1 - x0));
        }
        if (a.clip01) v = fminf(fmaxf(v, 0.f), 1.f);
        a.dst[c * onp + p] = v;
    }
}

// =================================================================================================
// Temporal stage: front end fused with the causal FIR  (cvvdp_metric.py:453-561)
// Every input frame is read and EOTF-ed exactly once per block (+ fl-1 history frames at the start of
// the block).  Four kernels; the host picks (cvvdp_api.cu, run_block):
//   k_temporal_2s  two pixels per thread (fp32x2), rolled front end + unrolled symmetric FIR with the ring in
//                  registers: dense, 16-byte aligned planes of whole 64-pixel segments (planar, channel-interleaved
//                  or planar YUV), 1..17 symmetric taps (images, 8..64 fps)                                   [default]
//   k_temporal_sr  the same ownership with the ring in shared memory: 19..73 taps (frame rates above 64 fps)
//   k_image_lut    images of 8-bit planes: the table front end alone
//   k_temporal     generic shared-memory ring, one pixel per thread: any layout (strided / permuted views, planar YUV
//                  of other widths), any filter
// (round 1 carried three more generations -- one pixel per thread with direct loads, cp.async-staged, packed
// but fully unrolled -- all superseded by the two-stage kernel and removed; measurements in DESIGN.md.)
// =================================================================================================
#define CVVDP_MEAN_SLOTS 64
struct TemporalArgs {
    ClipView clip[2];
    DisplayDev dd;
    int dtype, cin;
    int B, H, W;
    int F_total, f0, f1, fl, padding;
    YuvDev yuv;
    float4 *out;  // level 0: [B][n][2][H*W]
    int *flags;   // [0..2]: warps that saw out-of-range / NaN / Inf input values (null: no validation)
    int inter;    // two-stage / shared-ring kernels: 1 = channel-interleaved pixels (HWC frames: pixel stride 3, channel stride 1)
    float *mean0; // [CVVDP_MEAN_SLOTS] partial sums (a CTA / warp adds to slot index & (SLOTS-1)) over the pixels of clip frame 0 of the TEST video of its achromatic DKL channel (video_source.py:64-71)
    float taps[4][CVVDP_MAX_FILTER_LEN];  // taps[c][k] multiplies frame f-(fl-1)+k (= F_c flipped, l.556)
};

#define CVVDP_TEMPORAL_THREADS 256
__global__ void __launch_bounds__(CVVDP_TEMPORAL_THREADS) k_temporal(const TemporalArgs a) {
    CVVDP_DYN_SMEM(smem_raw);
    float *ring = reinterpret_cast<float *>(smem_raw);  // [fl][3][threads]
    const int tid = threadIdx.x;
    const long long npix = (long long)a.H * a.W;
    const long long p_raw = (long long)blockIdx.x * CVVDP_TEMPORAL_THREADS + tid;
    const int b = blockIdx.y >> 1, v = blockIdx.y & 1;
    // threads past the last pixel shadow it (no stores): every warp stays whole for the shuffles at the end
    const bool valid = p_raw < npix;
    const long long p = valid ? p_raw : npix - 1;
    const int y = (int)(p / a.W), x = (int)(p - (long long)y * a.W);
    const ClipView &cv = a.clip[v];
    const long long base = b * cv.s[0] + y * cv.s[3] + x * cv.s[4];
    const int n = a.f1 - a.f0, fl = a.fl;
    float4 *out = a.out + ((long long)b * n * 2 + v) * npix + p;
    int slot = 0, last_s = -1;
    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
    unsigned vbits = 0u;
    for (int t = a.f0 - (fl - 1); t < a.f1; ++t) {
        int s = t;
        if (s < 0) s = (a.padding == CVVDP_PAD_REPLICATE) ? 0 : symmetric_frame_index(s, a.F_total);
        if (s != last_s) {
            pixel_to_dkl(cv, base, frame_slot(cv, s), a.cin, a.dtype, a.dd, d0, d1, d2, &a.yuv, b, y, x, &vbits);
            last_s = s;
        }
        if (t == 0 && v == 0 && a.mean0 != nullptr) {  // uniform; one atomic per CTA (per thread they serialise: 3 ms per 1080p image)
            __shared__ float s_part[CVVDP_TEMPORAL_THREADS / 32];
            float msum = valid ? d0 : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) msum += __shfl_xor_sync(0xffffffffu, msum, o);
            if ((tid & 31) == 0) s_part[tid >> 5] = msum;
            __syncthreads();
            if (tid == 0) {
                float tot = 0.f;
                for (int w = 0; w < CVVDP_TEMPORAL_THREADS / 32; ++w) tot += s_part[w];
                atomicAdd(a.mean0 + (blockIdx.x & (CVVDP_MEAN_SLOTS - 1)), tot);
            }
        }
        ring[(slot * 3 + 0) * CVVDP_TEMPORAL_THREADS + tid] = d0;
        ring[(slot * 3 + 1) * CVVDP_TEMPORAL_THREADS + tid] = d1;
        ring[(slot * 3 + 2) * CVVDP_TEMPORAL_THREADS + tid] = d2;
        if (t >= a.f0) {
            float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
            int ks = slot + 1 == fl ? 0 : slot + 1;  // oldest frame in the ring
            for (int k = 0; k < fl; ++k) {
                const float x0 = ring[(ks * 3 + 0) * CVVDP_TEMPORAL_THREADS + tid];
                const float x1 = ring[(ks * 3 + 1) * CVVDP_TEMPORAL_THREADS + tid];
                const float x2 = ring[(ks * 3 + 2) * CVVDP_TEMPORAL_THREADS + tid];
                o0 = fmaf(a.taps[0][k], x0, o0);
                o1 = fmaf(a.taps[1][k], x1, o1);
                o2 = fmaf(a.taps[2][k], x2, o2);
                o3 = fmaf(a.taps[3][k], x0, o3);  // transient channel filters the achromatic plane (l.557)
                ks = ks + 1 == fl ? 0 : ks + 1;
            }
            if (valid) out[(long long)(t - a.f0) * 2 * npix] = make_float4(o0, o1, o2, o3);
        }
        slot = slot + 1 == fl ? 0 : slot + 1;
    }
    publish_input_bits(vbits, a.flags);
}

__device__ __forceinline__ float bits_as_float(unsigned b) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

__device__ __forceinline__ unsigned float_as_bits(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    unsigned b;
    memcpy(&b, &f, 4);
    return b;
#endif
}

__device__ __forceinline__ int temporal_source_frame(const TemporalArgs &a, int t) {
    if (t >= 0) return t;
    return (a.padding == CVVDP_PAD_REPLICATE) ? 0 : symmetric_frame_index(t, a.F_total);
}

// Pre-filtered sources (cvvdp_metric.py:470-488): the video source already delivers the four temporal channels
// (A-sust, RG, YV, A-trans: colour space 'DKLd65_trans'), so the FIR is bypassed and the stage only repacks the
// planar fp32 [B,4,F,H,W] frames into the float4 pixels of level 0.
struct PackArgs {
    ClipView clip[2];
    int B, H, W, f0, f1;
    float4 *out;  // level 0: [B][n][2][H*W]
};
__global__ void __launch_bounds__(256) k_pack_level0(const PackArgs a) {
    const long long npix = (long long)a.H * a.W;
    const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
    if (p >= npix) return;
    const int n = a.f1 - a.f0;
    const int v = blockIdx.y & 1, bf = blockIdx.y >> 1, b = bf / n, f = bf - b * n;
    const ClipView &cv = a.clip[v];
    const int y = (int)(p / a.W), x = (int)(p - (long long)y * a.W);
    const float *src = (const float *)cv.data + b * cv.s[0] + (long long)frame_slot(cv, a.f0 + f) * cv.s[2] + y * cv.s[3] + x * cv.s[4];
    a.out[((long long)(b * n + f) * 2 + v) * npix + p] = make_float4(src[0], src[cv.s[1]], src[2 * cv.s[1]], src[3 * cv.s[1]]);
}

// Two-stage packed temporal kernel.  A thread owns TWO pixels of its warp's 64-pixel segment (lane l: pixels
// l and l+32, so both 128-bit stores of a warp cover 512 contiguous bytes) as the halves of fp32x2 registers;
// the ring of the last FL frames lives in registers.  Unrolling the whole time loop by one ring period
// (round 1's first packed kernel) gave ~60 KB of straight-line code, twice the SM's 32 KB L1.5 instruction
// cache: ncu showed it waiting for instructions 2.4 cycles per issue.  Here only the FIR is unrolled.  Time advances in chunks of G = (FL+1)/2 frames:
//   stage 1 (a rolled loop, one copy of the code): raw values of the chunk (already in shared memory,
//            copied with cp.async during the previous chunk) -> EOTF -> DKL -> a per-thread slab of
//            shared memory (each thread reads back only what it wrote: no synchronisation);
//   stage 2 (unrolled over one ring period = two chunks): 3 LDS.64 bring a frame into its static ring
//            slot, 4*FL FFMA2, two 128-bit stores.
// The unrolled part shrinks to ~25 KB, the front end (including the generic per-pixel EOTF switch of
// the non-table variant) exists once, and the warps of a CTA never synchronise with each other.
// Per-lane arithmetic (EOTF table or eotf_forward, matrix order, tap order) as in the generic kernel.
// One byte from shared memory straight into a 32-bit register (the compiler otherwise packs pairs of
// 8-bit loads into 16-bit halves and unpacks them again).
__device__ __forceinline__ unsigned lds_u8(const unsigned char *p) {
#ifdef __CUDA_ARCH__
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
#else
    return *p;
#endif
}
// Two pixels (a, b) -> DKL, as the halves of fp32x2 values.  Per lane the same operations in the same
// order as bits_to_dkl (v1*M1, then fma with v0*M0, then fma with v2*M2), two lanes per instruction.
template <bool USE_LUT, int DT = -1, int E = -1>  // DT / E >= 0: dtype / EOTF fixed at compile time (three channels)
__device__ __forceinline__ void bits_to_dkl2(const TemporalArgs &a, const float *lut, const unsigned ba[3], const unsigned bb[3],
                                             float2 &d0, float2 &d1, float2 &d2, unsigned &vbits) {
    float va[3], vb[3];
    if (USE_LUT) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            va[i] = lut[ba[i]];
            vb[i] = lut[bb[i]];
        }
    } else {
        const int dtype = DT >= 0 ? DT : a.dtype;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            switch (dtype) {
                case CVVDP_DTYPE_U8: va[i] = (float)ba[i] / 255.0f; vb[i] = (float)bb[i] / 255.0f; break;
                case CVVDP_DTYPE_U16: va[i] = (float)ba[i] * (1.0f / 65535.0f); vb[i] = (float)bb[i] * (1.0f / 65535.0f); break;
                case CVVDP_DTYPE_F16: va[i] = half_bits_to_float((unsigned short)ba[i]); vb[i] = half_bits_to_float((unsigned short)bb[i]); break;
                default: va[i] = bits_as_float(ba[i]); vb[i] = bits_as_float(bb[i]);
            }
        }
        if (dtype >= CVVDP_DTYPE_F16) {  // integer code values cannot be out of range, NaN or Inf
            const int eotf = E >= 0 ? E : a.dd.eotf;
            const bool rng = eotf != CVVDP_EOTF_LINEAR && eotf != CVVDP_EOTF_NONE;
            // As unsigned integers, the floats of [0, 1] are exactly the patterns <= 0x3f800000 (and -0); with the sign
            // masked off, the finite ones are those <= 0x7f7fffff.  One integer max per value finds out whether this
            // frame needs the exact classification at all.
            unsigned worst = 0u;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const unsigned xa = float_as_bits(va[i]), xb = float_as_bits(vb[i]);
                worst = max(worst, max(rng ? xa : (xa & 0x7fffffffu), rng ? xb : (xb & 0x7fffffffu)));
            }
            if (worst > (rng ? 0x3f800000u : 0x7f7fffffu)) {
#pragma unroll
                for (int i = 0; i < 3; ++i) vbits |= input_bits(va[i], rng) | input_bits(vb[i], rng);
            }
        }
        if (E >= 0) {
            eotf_forward_c<E>(va, a.dd);
            eotf_forward_c<E>(vb, a.dd);
        } else {
            eotf_forward(va, a.cin, a.dd);
            eotf_forward(vb, a.cin, a.dd);
        }
    }
    if (E >= 0 || a.cin == 3) {
        const float2 v0 = make_float2(va[0], vb[0]), v1 = make_float2(va[1], vb[1]), v2 = make_float2(va[2], vb[2]);
        d0 = fma2(v2, bc2(a.dd.M[2]), fma2(v0, bc2(a.dd.M[0]), mul2(v1, bc2(a.dd.M[1]))));
        d1 = fma2(v2, bc2(a.dd.M[5]), fma2(v0, bc2(a.dd.M[3]), mul2(v1, bc2(a.dd.M[4]))));
        d2 = fma2(v2, bc2(a.dd.M[8]), fma2(v0, bc2(a.dd.M[6]), mul2(v1, bc2(a.dd.M[7]))));
    } else {
        d0 = d1 = d2 = make_float2(va[0], vb[0]);
    }
}

#define CVVDP_T2S_THREADS 128
#ifndef CVVDP_T2S_UNROLL_LUT
#define CVVDP_T2S_UNROLL_LUT 3   // frames of the stage-1 loop in flight: table variant
#endif
#ifndef CVVDP_T2S_UNROLL_SPEC
#define CVVDP_T2S_UNROLL_SPEC 2  // ... compile-time (dtype, EOTF) bodies
#endif
constexpr int kT2sUnrollLut = CVVDP_T2S_UNROLL_LUT, kT2sUnrollSpec = CVVDP_T2S_UNROLL_SPEC;
template <int FL>
struct T2SGeom {
    static constexpr int RP = FL + 1;   // ring period (even)
    static constexpr int G = RP / 2;    // frames per chunk
};
// dynamic shared memory: [warps][G][3 (planar YUV: 5)][64*esz] raw bytes, then [G][3][threads] float2
__host__ __device__ inline size_t t2s_smem_bytes(int fl, int esz, int rows = 3) {  // rows: 3 channels, or Y + 2 x 2 chroma rows
    const int G = (fl + 1) / 2;
    return (size_t)(CVVDP_T2S_THREADS / 32) * G * rows * 64 * esz + (size_t)G * 3 * CVVDP_T2S_THREADS * 8;
}
// SRC: where the frames come from -- 0: dense planes of any dtype (raw stage filled by cp.async), 1: 8-bit planes through
// the 256-entry EOTF table, 2: planar YUV frames read straight from global memory (the chroma taps of neighbouring
// pixels overlap, so L1 serves most of them; rows must be whole 64-pixel segments).
#define CVVDP_T2S_ANY 0
#define CVVDP_T2S_LUT 1
#define CVVDP_T2S_YUV 2
template <int FL, int SRC>
__global__ void __launch_bounds__(CVVDP_T2S_THREADS, 3) k_temporal_2s(const __grid_constant__ TemporalArgs a) {
    constexpr bool USE_LUT = SRC == CVVDP_T2S_LUT;
    constexpr int RP = T2SGeom<FL>::RP, G = T2SGeom<FL>::G;
    __shared__ float s_lut[256];
    CVVDP_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (USE_LUT) {
        for (int i = tid; i < 256; i += CVVDP_T2S_THREADS) {
            float v[3] = {(float)i / 255.0f, 0.f, 0.f};
            eotf_forward_n<1>(v, a.dd);
            s_lut[i] = v[0];
        }
        __syncthreads();
    }
    const long long npix = (long long)a.H * a.W;
    const long long wp = ((long long)blockIdx.x * (CVVDP_T2S_THREADS / 32) + warp) * 64;  // first pixel of the warp
    const int b = blockIdx.y >> 1, v = blockIdx.y & 1;
    if (wp >= npix) return;  // whole 64-pixel segments only (npix % 64 == 0); no block-level barrier below
    const ClipView &cv = a.clip[v];
    constexpr bool YUV = SRC == CVVDP_T2S_YUV;
    const int esz = USE_LUT ? 1 : (a.dtype == CVVDP_DTYPE_F32 ? 4 : (a.dtype == CVVDP_DTYPE_U8 ? 1 : 2));
    const int row_bytes = 64 * esz;          // one (channel, frame) segment of this warp
    // slot of one frame in the raw stage: three channel segments (cin == 1 uses the first), or -- planar YUV -- the luma
    // segment and two rows of 64 samples of each chroma plane around it
    const int frame_bytes = (YUV ? 5 : 3) * row_bytes;
    const int cpc = row_bytes / 16;          // 16-byte pieces per segment
    const int ppf = YUV ? 5 * cpc : a.cin * cpc;  // pieces per frame
    unsigned char *raw = smem_raw + (size_t)warp * G * frame_bytes;
    float2 *dkl = reinterpret_cast<float2 *>(smem_raw + (size_t)(CVVDP_T2S_THREADS / 32) * G * frame_bytes) + tid;
    // Channel-interleaved frames (HWC): the warp's 64 pixels are 3 * 64 consecutive elements -- for the copies that is
    // three "channel segments" 64 elements apart; only the element a lane picks out of the stage differs (below).
    const bool inter = !YUV && a.inter != 0;
    const long long fstride = cv.s[2] * esz, cstride = inter ? row_bytes : cv.s[1] * esz;
    const unsigned char *wsrc = (const unsigned char *)cv.data + (b * cv.s[0] + (inter ? 3 * wp : wp)) * esz;
    const int n = a.f1 - a.f0;
    const int NI = (FL - 1) + n;  // iterations: FL-1 warm-up frames (temporal padding before frame 0), then the block
    // the thread's two pixels: lane and lane + 32 of the segment, or -- planar YUV -- the horizontal neighbours 2 lane and
    // 2 lane + 1, which share their chroma samples
    constexpr int PB = SRC == CVVDP_T2S_YUV ? 1 : 32;
    float4 *outp = a.out + ((long long)b * n * 2 + v) * npix + wp + (SRC == CVVDP_T2S_YUV ? 2 * lane : lane);
    const long long ostep = 2 * npix;
    float2 r0[RP], r1[RP], r2[RP];
#pragma unroll
    for (int i = 0; i < RP; ++i) r0[i] = r1[i] = r2[i] = make_float2(0.f, 0.f);

    // cp.async copies of chunk `c` (iterations c*G .. c*G+G-1) into the raw stage; issued as soon as the
    // previous chunk is converted, so they land while the FIR of that chunk runs.  Piece p = lane + 32 r of
    // a chunk is (frame g, channel, 16-byte part); (g, rem) advance incrementally from round to round.
    const int q32 = 32 / ppf, r32 = 32 - q32 * ppf;  // 32 = q32 * ppf + r32
    const int g_first = lane / ppf, rem_first = lane - g_first * ppf;
    const int cpc_shift = esz == 1 ? 2 : (esz == 2 ? 3 : 4);
    // Planar YUV (video_source_yuv.py:153-233): the warp's 64 pixels lie in one row y (rows are whole segments).  Staged per
    // frame: the 64 luma samples and, of each chroma plane, rows j0 and j1 (yuv_chroma_taps) from sample yuv_cs on --
    // one 16-byte piece to the left of the segment's first chroma sample, so that the left neighbour is there and every
    // piece stays 16-byte aligned; pieces that start outside the row are skipped (the taps clamp to the row).
    int yuv_j0 = 0, yuv_j1 = 0, yuv_cs = 0, yuv_cw = 0, yuv_ypix = 0, yuv_uvpix = 0;  // (a frame has fewer than 2^31 samples)
    float yuv_ly = 0.f;
    if (YUV) {
        yuv_cw = a.yuv.chroma == 444 ? a.W : a.W / 2;
        yuv_ypix = (int)npix;
        yuv_uvpix = yuv_cw * (a.yuv.chroma == 420 ? a.H / 2 : a.H);
        const int y = (int)(wp / a.W), x0 = (int)(wp - (long long)y * a.W);
        int i0, i1;
        float lx_unused;
        yuv_chroma_taps(a.yuv, y, x0, i0, i1, yuv_j0, yuv_j1, lx_unused, yuv_ly);
        yuv_cs = a.yuv.chroma == 444 ? x0 : x0 / 2 - 16 / esz;
    }
    auto issue_chunk = [&](int c) {
        int g = g_first, rem = rem_first;
        const int it0 = c * G;
        while (g < G) {
            const int it = it0 + g;
            if (it < NI) {
                const int t = a.f0 - (FL - 1) + it;
                const int slot = frame_slot(cv, t >= 0 ? t : temporal_source_frame(a, t));
                const int ch = rem >> cpc_shift, part = rem & (cpc - 1);
                if (!YUV || ch == 0) {
                    cp_async16(raw + g * frame_bytes + ch * row_bytes + part * 16, wsrc + (long long)slot * fstride + ch * cstride + part * 16);
                } else {  // segment 1 + 2 * plane + row
                    const int ps = yuv_cs + part * (16 / esz);
                    if (ps >= 0 && ps < yuv_cw) {
                        const long long so = yuv_ypix + ((ch - 1) >> 1) * (long long)yuv_uvpix + (long long)(((ch - 1) & 1) ? yuv_j1 : yuv_j0) * yuv_cw + ps;
                        cp_async16(raw + g * frame_bytes + ch * row_bytes + part * 16, wsrc + (long long)slot * fstride + (so - wp) * esz);
                    }
                }
            }
            g += q32;
            rem += r32;
            if (rem >= ppf) {
                rem -= ppf;
                ++g;
            }
        }
        cp_async_commit();
    };
    // stage 1: raw -> DKL for the G frames of the chunk in the raw stage (rolled; three frames in flight)
    unsigned vbits = 0u;  // input validity bits seen by this thread (floating-point clips only)
    float msum = 0.f;     // achromatic DKL sum of this thread's two pixels of clip frame 0 (test video)
    const int it_zero = v == 0 && a.mean0 != nullptr ? FL - 1 - a.f0 : -1;  // iteration that holds clip frame 0
    // One body per (dtype, EOTF) pair that matters for throughput -- fp32 / fp16 / uint16 clips on an sRGB, PQ or linear
    // display, three channels -- with both fixed at compile time; everything else (and the table variant, which has
    // nothing left to specialise) takes the generic body with its run-time switches.
    auto convert_body = [&](auto dt_c, auto eotf_c, auto esz_c, int it0) {
        constexpr int DT = decltype(dt_c)::value, E = decltype(eotf_c)::value, ESZ = decltype(esz_c)::value;
        const int eszv = ESZ > 0 ? ESZ : esz;
        // three frames in flight for the table variant (a dozen instructions per frame), two for a specialised body; the
        // generic per-pixel EOTF is long enough to hide its own latencies and must not be replicated
#pragma unroll(USE_LUT ? kT2sUnrollLut : (DT >= 0 ? kT2sUnrollSpec : 1))
        for (int g = 0; g < G; ++g) {
            const unsigned char *q = raw + g * frame_bytes;
            unsigned ba[3], bb[3];
            if (!inter) {  // uniform; planar: the offsets of a frame's channels are immediates of the unrolled body
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const unsigned char *qc = q + ((DT >= 0 || a.cin == 3) ? ch : 0) * row_bytes;
                    if (eszv == 1) {
                        ba[ch] = lds_u8(qc + lane);
                        bb[ch] = lds_u8(qc + lane + 32);
                    } else if (eszv == 2) {
                        ba[ch] = ((const unsigned short *)qc)[lane];
                        bb[ch] = ((const unsigned short *)qc)[lane + 32];
                    } else {
                        ba[ch] = ((const unsigned *)qc)[lane];
                        bb[ch] = ((const unsigned *)qc)[lane + 32];
                    }
                }
            } else {  // channel-interleaved: elements 3 lane + ch and 3 (lane + 32) + ch of the frame's 192
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    if (eszv == 1) {
                        ba[ch] = lds_u8(q + 3 * lane + ch);
                        bb[ch] = lds_u8(q + 3 * lane + 96 + ch);
                    } else if (eszv == 2) {
                        ba[ch] = ((const unsigned short *)q)[3 * lane + ch];
                        bb[ch] = ((const unsigned short *)q)[3 * lane + 96 + ch];
                    } else {
                        ba[ch] = ((const unsigned *)q)[3 * lane + ch];
                        bb[ch] = ((const unsigned *)q)[3 * lane + 96 + ch];
                    }
                }
            }
            float2 d0, d1, d2;
            bits_to_dkl2<USE_LUT, DT, E>(a, s_lut, ba, bb, d0, d1, d2, vbits);
            if (it0 + g == it_zero) msum = d0.x + d0.y;  // uniform
            dkl[(g * 3 + 0) * CVVDP_T2S_THREADS] = d0;
            dkl[(g * 3 + 1) * CVVDP_T2S_THREADS] = d1;
            dkl[(g * 3 + 2) * CVVDP_T2S_THREADS] = d2;
        }
    };
    // Planar YUV front end: the thread's pixels (y, x) and (y, x + 1), x even, need the chroma columns L, M, R of rows j0, j1
    // of each plane -- six samples instead of eight: with subsampled chroma (k = x / 2) L = max(k-1, 0), M = k,
    // R = min(k+1, cw-1), pixel x blends (L, M) with weight 0.75 (0 at k = 0, where torch clamps the source coordinate) and
    // pixel x + 1 blends (M, R) with 0.25; 4:4:4 takes L = x, M = R = x + 1 with weights 0.  The limited-range clip of a
    // chroma sample is done as saturate(c + 0.5) and the 0.5 taken off after the blend (the weights add up to one).
    int yuv_iL = 0, yuv_iM = 0, yuv_iR = 0;  // sample indices in a staged chroma row
    float2 yuv_lx = make_float2(0.f, 0.f);
    if (YUV) {
        const int x = (int)(wp % a.W) + 2 * lane;
        if (a.yuv.chroma == 444) {
            yuv_iL = x - yuv_cs;
            yuv_iM = yuv_iR = x + 1 - yuv_cs;
        } else {
            const int k = x >> 1;
            yuv_iL = max(k - 1, 0) - yuv_cs;
            yuv_iM = k - yuv_cs;
            yuv_iR = min(k + 1, yuv_cw - 1) - yuv_cs;
            yuv_lx = make_float2(k == 0 ? 0.f : 0.75f, 0.25f);
        }
    }
    auto convert_yuv = [&](auto dt_c, auto eotf_c, int it0) {
        constexpr int DT = decltype(dt_c)::value, E = decltype(eotf_c)::value;
        typedef typename std::conditional<DT == CVVDP_DTYPE_U8, unsigned char, unsigned short>::type S;
        const YuvDev &yu = a.yuv;
        const float c_off = 0.5f - yu.co;
        const float2 wx1 = yuv_lx, wx0 = make_float2(1.f - yuv_lx.x, 1.f - yuv_lx.y);
        const int row1 = yuv_j1 != yuv_j0 ? 64 : 0;  // samples from row j0 to row j1 in the stage
#pragma unroll(FL >= 15 ? 1 : 2)  // (the 17-tap ring leaves no registers for a second frame in flight)
        for (int g = 0; g < G; ++g) {
            const S *q = (const S *)(raw + g * frame_bytes);
            const float Ya = __saturatef(fmaf(yu.yw, (float)q[2 * lane], -yu.yo)), Yb = __saturatef(fmaf(yu.yw, (float)q[2 * lane + 1], -yu.yo));
            float2 uv[2];
#pragma unroll
            for (int pl = 0; pl < 2; ++pl) {
                const S *qc = q + 64 + 128 * pl;
                const float cL0 = __saturatef(fmaf(yu.cw, (float)qc[yuv_iL], c_off));
                const float cM0 = __saturatef(fmaf(yu.cw, (float)qc[yuv_iM], c_off));
                const float cR0 = __saturatef(fmaf(yu.cw, (float)qc[yuv_iR], c_off));
                const float cL1 = __saturatef(fmaf(yu.cw, (float)qc[row1 + yuv_iL], c_off));
                const float cM1 = __saturatef(fmaf(yu.cw, (float)qc[row1 + yuv_iM], c_off));
                const float cR1 = __saturatef(fmaf(yu.cw, (float)qc[row1 + yuv_iR], c_off));
                const float2 top = fma2(wx0, make_float2(cL0, cM0), mul2(wx1, make_float2(cM0, cR0)));
                const float2 bot = fma2(wx0, make_float2(cL1, cM1), mul2(wx1, make_float2(cM1, cR1)));
                uv[pl] = add2(fma2(bc2(1.f - yuv_ly), top, mul2(bc2(yuv_ly), bot)), bc2(-0.5f));
            }
            float ra[3], rb[3];
            ra[0] = __saturatef(fmaf(yu.m_rv, uv[1].x, Ya));
            ra[1] = __saturatef(fmaf(yu.m_gv, uv[1].x, fmaf(yu.m_gu, uv[0].x, Ya)));
            ra[2] = __saturatef(fmaf(yu.m_bu, uv[0].x, Ya));
            rb[0] = __saturatef(fmaf(yu.m_rv, uv[1].y, Yb));
            rb[1] = __saturatef(fmaf(yu.m_gv, uv[1].y, fmaf(yu.m_gu, uv[0].y, Yb)));
            rb[2] = __saturatef(fmaf(yu.m_bu, uv[0].y, Yb));
            if (E >= 0) {
                eotf_forward_c<E, true>(ra, a.dd);
                eotf_forward_c<E, true>(rb, a.dd);
            } else {
                eotf_forward_n<3>(ra, a.dd);
                eotf_forward_n<3>(rb, a.dd);
            }
            const float2 v0 = make_float2(ra[0], rb[0]), v1 = make_float2(ra[1], rb[1]), v2 = make_float2(ra[2], rb[2]);
            const float2 d0 = fma2(v2, bc2(a.dd.M[2]), fma2(v0, bc2(a.dd.M[0]), mul2(v1, bc2(a.dd.M[1]))));
            const float2 d1 = fma2(v2, bc2(a.dd.M[5]), fma2(v0, bc2(a.dd.M[3]), mul2(v1, bc2(a.dd.M[4]))));
            const float2 d2 = fma2(v2, bc2(a.dd.M[8]), fma2(v0, bc2(a.dd.M[6]), mul2(v1, bc2(a.dd.M[7]))));
            if (it0 + g == it_zero) msum = d0.x + d0.y;  // uniform
            dkl[(g * 3 + 0) * CVVDP_T2S_THREADS] = d0;
            dkl[(g * 3 + 1) * CVVDP_T2S_THREADS] = d1;
            dkl[(g * 3 + 2) * CVVDP_T2S_THREADS] = d2;
        }
    };
    // 0: generic; 1 + 3 * k + e: k = fp32, fp16, uint16 and e = sRGB, PQ, linear
    int conv_mode = 0;
    if (!USE_LUT && a.cin == 3) {
        const int k = a.dtype == CVVDP_DTYPE_F32 ? 0 : (a.dtype == CVVDP_DTYPE_F16 ? 1 : (a.dtype == CVVDP_DTYPE_U16 ? 2 : -1));
        const int e = a.dd.eotf == CVVDP_EOTF_SRGB ? 0 : (a.dd.eotf == CVVDP_EOTF_PQ ? 1 : (a.dd.eotf == CVVDP_EOTF_LINEAR ? 2 : -1));
        if (k >= 0 && e >= 0) conv_mode = 1 + 3 * k + e;
    }
    auto convert_chunk = [&](int it0) {
        using std::integral_constant;
#define CVVDP_CONV_CASE(MODE, DTV, EV, ESZV)                                                                                    \
    case MODE:                                                                                                                  \
        convert_body(integral_constant<int, DTV>{}, integral_constant<int, EV>{}, integral_constant<int, ESZV>{}, it0);         \
        break;
        if (USE_LUT) {
            convert_body(integral_constant<int, -1>{}, integral_constant<int, -1>{}, integral_constant<int, 1>{}, it0);
            return;
        }
        if constexpr (SRC == CVVDP_T2S_YUV) {  // 8- or 16-bit samples x (sRGB, PQ, any other EOTF at run time)
            const int e = a.dd.eotf;
            if (a.dtype == CVVDP_DTYPE_U8) {
                if (e == CVVDP_EOTF_SRGB) convert_yuv(integral_constant<int, CVVDP_DTYPE_U8>{}, integral_constant<int, CVVDP_EOTF_SRGB>{}, it0);
                else if (e == CVVDP_EOTF_PQ) convert_yuv(integral_constant<int, CVVDP_DTYPE_U8>{}, integral_constant<int, CVVDP_EOTF_PQ>{}, it0);
                else convert_yuv(integral_constant<int, CVVDP_DTYPE_U8>{}, integral_constant<int, -1>{}, it0);
            } else {
                if (e == CVVDP_EOTF_SRGB) convert_yuv(integral_constant<int, CVVDP_DTYPE_U16>{}, integral_constant<int, CVVDP_EOTF_SRGB>{}, it0);
                else if (e == CVVDP_EOTF_PQ) convert_yuv(integral_constant<int, CVVDP_DTYPE_U16>{}, integral_constant<int, CVVDP_EOTF_PQ>{}, it0);
                else convert_yuv(integral_constant<int, CVVDP_DTYPE_U16>{}, integral_constant<int, -1>{}, it0);
            }
            return;
        }
        if constexpr (SRC == CVVDP_T2S_ANY) {
            switch (conv_mode) {  // uniform over the grid
            CVVDP_CONV_CASE(1, CVVDP_DTYPE_F32, CVVDP_EOTF_SRGB, 4)
            CVVDP_CONV_CASE(2, CVVDP_DTYPE_F32, CVVDP_EOTF_PQ, 4)
            CVVDP_CONV_CASE(3, CVVDP_DTYPE_F32, CVVDP_EOTF_LINEAR, 4)
            CVVDP_CONV_CASE(4, CVVDP_DTYPE_F16, CVVDP_EOTF_SRGB, 2)
            CVVDP_CONV_CASE(5, CVVDP_DTYPE_F16, CVVDP_EOTF_PQ, 2)
            CVVDP_CONV_CASE(6, CVVDP_DTYPE_F16, CVVDP_EOTF_LINEAR, 2)
            CVVDP_CONV_CASE(7, CVVDP_DTYPE_U16, CVVDP_EOTF_SRGB, 2)
            CVVDP_CONV_CASE(8, CVVDP_DTYPE_U16, CVVDP_EOTF_PQ, 2)
            CVVDP_CONV_CASE(9, CVVDP_DTYPE_U16, CVVDP_EOTF_LINEAR, 2)
            default: convert_body(integral_constant<int, -1>{}, integral_constant<int, -1>{}, integral_constant<int, 0>{}, it0);
            }
        }
#undef CVVDP_CONV_CASE
    };
    issue_chunk(0);
    int it = 0;
    for (int c = 0; it < NI; c += 2) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {  // the two chunks of one ring period
            cp_async_wait_all();
            __syncwarp();          // the chunk's raw values (copied by all lanes) are visible
            convert_chunk((c + h) * G);
            __syncwarp();          // every lane is done with the raw stage before it is refilled
            issue_chunk(c + h + 1);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int s = h * G + g;  // static ring slot: it % RP
                if (it < NI) {            // uniform
                    r0[s] = dkl[(g * 3 + 0) * CVVDP_T2S_THREADS];
                    r1[s] = dkl[(g * 3 + 1) * CVVDP_T2S_THREADS];
                    r2[s] = dkl[(g * 3 + 2) * CVVDP_T2S_THREADS];
                    if (it >= FL - 1) {   // uniform
                        // the filters are exactly symmetric (the host checks, else the generic kernel runs): frames at
                        // mirrored taps are added first, which also lets A-sust and A-trans share the sums
                        // (splitting each sum into two chains for more ILP was tried: +29 instructions per pixel of
                        // rematerialised addresses under register pressure, 13 % slower)
                        float2 o0 = make_float2(0.f, 0.f), o1 = o0, o2 = o0, o3 = o0;
#pragma unroll
                        for (int k = 0; k < FL / 2; ++k) {  // tap k <-> frame t-(FL-1)+k <-> slot (s-(FL-1)+k) mod RP
                            const int sa = (s + RP - (FL - 1) + k) % RP, sb = (s + RP - k) % RP;
                            const float2 p0 = add2(r0[sa], r0[sb]), p1 = add2(r1[sa], r1[sb]), p2 = add2(r2[sa], r2[sb]);
                            o0 = fma2(bc2(a.taps[0][k]), p0, o0);
                            o1 = fma2(bc2(a.taps[1][k]), p1, o1);
                            o2 = fma2(bc2(a.taps[2][k]), p2, o2);
                            o3 = fma2(bc2(a.taps[3][k]), p0, o3);
                        }
                        {
                            constexpr int k = FL / 2;
                            const int sl = (s + RP - (FL - 1) + k) % RP;
                            o0 = fma2(bc2(a.taps[0][k]), r0[sl], o0);
                            o1 = fma2(bc2(a.taps[1][k]), r1[sl], o1);
                            o2 = fma2(bc2(a.taps[2][k]), r2[sl], o2);
                            o3 = fma2(bc2(a.taps[3][k]), r0[sl], o3);
                        }
                        outp[0] = make_float4(o0.x, o1.x, o2.x, o3.x);
                        outp[PB] = make_float4(o0.y, o1.y, o2.y, o3.y);
                        outp += ostep;
                    }
                }
                ++it;
            }
        }
    }
    cp_async_wait_all();
    if (!USE_LUT) publish_input_bits(vbits, a.flags);
    if (it_zero >= 0 && it_zero < NI) {  // uniform per CTA
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) msum += __shfl_xor_sync(0xffffffffu, msum, o);
        // (the slot comes from the block index alone: keeping `warp` alive to this point tipped the register allocator of
        // the two-stage kernel into rematerialising thread and block indices inside the frame loop -- S2R / LEA / IMAD per
        // frame, 8.3 -> 9.8 ms for the 8-bit and 14.5 -> 18.9 ms for the fp32 4K clip; profiles/r02_ab_temporal_remat.txt)
        if (lane == 0) atomicAdd(a.mean0 + (blockIdx.x & (CVVDP_MEAN_SLOTS - 1)), msum);
    }
}

// =================================================================================================
// Images of 8-bit planes (the reference's everyday case: predict(img_test, img_ref, 'HWC'), batches of them): no time
// axis, so the whole temporal stage is unpack -> EOTF table -> DKL -> level 0.  A thread converts four consecutive
// pixels per step (three 32-bit loads, twelve table look-ups, four 128-bit stores); CTAs stride over the plane so that
// the 256-entry table is built once per CTA.  Planar and channel-interleaved frames; dense, 16-byte aligned planes of
// whole 64-pixel segments (the host checks), everything else goes to k_temporal_2s<1> / k_temporal.
// Per pixel the arithmetic of bits_to_dkl2 and of the one-tap FIR (tap x value), so the bits match those kernels.
// =================================================================================================
__global__ void __launch_bounds__(256) k_image_lut(const __grid_constant__ TemporalArgs a) {
    __shared__ float s_lut[256];
    __shared__ float s_part[8];
    const int tid = threadIdx.x;
    {
        float v[3] = {(float)tid / 255.0f, 0.f, 0.f};
        eotf_forward_n<1>(v, a.dd);
        s_lut[tid] = v[0];
    }
    __syncthreads();
    const long long npix = (long long)a.H * a.W;
    const int b = blockIdx.y >> 1, v = blockIdx.y & 1;
    const ClipView &cv = a.clip[v];
    const bool inter = a.inter != 0;
    const unsigned char *src = (const unsigned char *)cv.data + b * cv.s[0] + (long long)frame_slot(cv, a.f0) * cv.s[2];
    float4 *out = a.out + ((long long)b * 2 + v) * npix;
    const float t0 = a.taps[0][0], t1 = a.taps[1][0], t2 = a.taps[2][0], t3 = a.taps[3][0];
    float msum = 0.f;
    for (long long p = ((long long)blockIdx.x * 256 + tid) * 4; p < npix; p += (long long)gridDim.x * 1024) {
        unsigned w0, w1, w2;  // planar: four pixels of each channel; interleaved: the twelve bytes of four pixels
        if (inter) {
            const unsigned *q = reinterpret_cast<const unsigned *>(src + 3 * p);
            w0 = __ldg(q);
            w1 = __ldg(q + 1);
            w2 = __ldg(q + 2);
        } else {
            w0 = __ldg(reinterpret_cast<const unsigned *>(src + p));
            w1 = a.cin == 3 ? __ldg(reinterpret_cast<const unsigned *>(src + cv.s[1] + p)) : w0;
            w2 = a.cin == 3 ? __ldg(reinterpret_cast<const unsigned *>(src + 2 * cv.s[1] + p)) : w0;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            unsigned c0, c1, c2;
            if (inter) {  // byte 3 i + ch of the twelve
                const unsigned long long lo = ((unsigned long long)w1 << 32) | w0;
                c0 = i < 3 ? (unsigned)(lo >> (24 * i)) & 0xffu : (w2 >> 8) & 0xffu;
                c1 = i < 2 ? (unsigned)(lo >> (24 * i + 8)) & 0xffu : (i == 2 ? (w1 >> 24) & 0xffu : (w2 >> 16) & 0xffu);
                c2 = i < 2 ? (unsigned)(lo >> (24 * i + 16)) & 0xffu : (i == 2 ? w2 & 0xffu : (w2 >> 24) & 0xffu);
            } else {
                c0 = (w0 >> (8 * i)) & 0xffu;
                c1 = (w1 >> (8 * i)) & 0xffu;
                c2 = (w2 >> (8 * i)) & 0xffu;
            }
            const float v0 = s_lut[c0], v1 = s_lut[c1], v2 = s_lut[c2];
            float d0, d1, d2;
            if (a.cin == 3) {  // the operation order of bits_to_dkl2
                d0 = fmaf(v2, a.dd.M[2], fmaf(v0, a.dd.M[0], v1 * a.dd.M[1]));
                d1 = fmaf(v2, a.dd.M[5], fmaf(v0, a.dd.M[3], v1 * a.dd.M[4]));
                d2 = fmaf(v2, a.dd.M[8], fmaf(v0, a.dd.M[6], v1 * a.dd.M[7]));
            } else {
                d0 = d1 = d2 = v0;
            }
            msum += d0;
            out[p + i] = make_float4(t0 * d0, t1 * d1, t2 * d2, t3 * d0);
        }
    }
    if (v == 0 && a.mean0 != nullptr && a.f0 == 0) {  // uniform per CTA
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) msum += __shfl_xor_sync(0xffffffffu, msum, o);
        if ((tid & 31) == 0) s_part[tid >> 5] = msum;
        __syncthreads();
        if (tid == 0) {
            float tot = 0.f;
            for (int w = 0; w < 8; ++w) tot += s_part[w];
            atomicAdd(a.mean0 + (blockIdx.x & (CVVDP_MEAN_SLOTS - 1)), tot);
        }
    }
}

// =================================================================================================
// Temporal stage for long filters (19 taps and more: frame rates above 64 fps; 120 fps = 31 taps).  The register ring
// of k_temporal_2s does not fit, so the ring lives in shared memory: [FL][3][threads] fp32x2, one slab per thread
// (nobody else reads it: no barriers).  Same pixel ownership (lane, lane + 32 of a 64-pixel segment) and the same
// symmetric-tap FIR (mirrored frames added first, A-sust and A-trans share the sums) as k_temporal_2s; the raw
// samples of the next frame are loaded before the FIR of the current one runs.  Dense planes only (the host checks);
// anything else stays with k_temporal.
// dynamic shared memory: float4 taps[FL/2 + 1] | float lut[256] (table variant) | float2 ring[FL][3][threads]
// =================================================================================================
#define CVVDP_TSR_THREADS 128
__host__ __device__ inline size_t tsr_smem_bytes(int fl, bool lut) {
    return (size_t)(fl / 2 + 1) * 16 + (lut ? 1024 : 0) + (size_t)(fl + 1) * 3 * CVVDP_TSR_THREADS * 8;
}
template <bool USE_LUT>
__global__ void __launch_bounds__(CVVDP_TSR_THREADS) k_temporal_sr(const __grid_constant__ TemporalArgs a) {
    CVVDP_DYN_SMEM(smem_raw);
    const int FL = a.fl, HP = FL / 2;  // HP mirrored pairs + the centre tap
    const int RS = FL + 1;             // ring slots: the filter support of two consecutive output frames
    float4 *s_taps = reinterpret_cast<float4 *>(smem_raw);
    float *s_lut = reinterpret_cast<float *>(smem_raw + (size_t)(HP + 1) * 16);
    float2 *ring = reinterpret_cast<float2 *>(smem_raw + (size_t)(HP + 1) * 16 + (USE_LUT ? 1024 : 0)) + threadIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k <= HP; k += CVVDP_TSR_THREADS) s_taps[k] = make_float4(a.taps[0][k], a.taps[1][k], a.taps[2][k], a.taps[3][k]);
    if (USE_LUT) {
        for (int i = tid; i < 256; i += CVVDP_TSR_THREADS) {
            float v[3] = {(float)i / 255.0f, 0.f, 0.f};
            eotf_forward_n<1>(v, a.dd);
            s_lut[i] = v[0];
        }
    }
    __syncthreads();
    const long long npix = (long long)a.H * a.W;
    const long long wp = ((long long)blockIdx.x * (CVVDP_TSR_THREADS / 32) + warp) * 64;
    const int b = blockIdx.y >> 1, v = blockIdx.y & 1;
    if (wp >= npix) return;  // whole segments only (npix % 64 == 0); no barrier below
    const ClipView &cv = a.clip[v];
    const int n = a.f1 - a.f0, NI = (FL - 1) + n;
    float4 *outp = a.out + ((long long)b * n * 2 + v) * npix + wp + lane;
    const long long ostep = 2 * npix;
    const int ps = a.inter ? 3 : 1;  // pixel stride: channel-interleaved (HWC) or planar frames
    const long long pbase = b * cv.s[0] + (wp + lane) * ps;  // element offset of the thread's first pixel in a frame
    auto load_raw = [&](int it, unsigned (&ra)[3], unsigned (&rb)[3]) {
        if (it >= NI) return;  // uniform
        const int t = a.f0 - (FL - 1) + it;
        const long long off = pbase + (long long)frame_slot(cv, t >= 0 ? t : temporal_source_frame(a, t)) * cv.s[2];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const long long o = off + (a.cin == 3 ? ch : 0) * cv.s[1];
            if (USE_LUT || a.dtype == CVVDP_DTYPE_U8) {
                ra[ch] = __ldg((const unsigned char *)cv.data + o);
                rb[ch] = __ldg((const unsigned char *)cv.data + o + 32 * ps);
            } else if (a.dtype == CVVDP_DTYPE_F32) {
                ra[ch] = __ldg((const unsigned *)cv.data + o);
                rb[ch] = __ldg((const unsigned *)cv.data + o + 32 * ps);
            } else {
                ra[ch] = __ldg((const unsigned short *)cv.data + o);
                rb[ch] = __ldg((const unsigned short *)cv.data + o + 32 * ps);
            }
        }
    };
    unsigned vbits = 0u;
    float msum = 0.f;
    const int it_zero = v == 0 && a.mean0 != nullptr ? FL - 1 - a.f0 : -1;
    auto convert = [&](int it, int slot, const unsigned (&ra)[3], const unsigned (&rb)[3]) {
        float2 d0, d1, d2;
        bits_to_dkl2<USE_LUT>(a, s_lut, ra, rb, d0, d1, d2, vbits);
        if (it == it_zero) msum = d0.x + d0.y;  // uniform
        ring[(slot * 3 + 0) * CVVDP_TSR_THREADS] = d0;
        ring[(slot * 3 + 1) * CVVDP_TSR_THREADS] = d1;
        ring[(slot * 3 + 2) * CVVDP_TSR_THREADS] = d2;
    };
    auto slot_add = [&](int s, int d) {  // (s + d) mod RS for 0 <= d <= 2
        s += d;
        return s >= RS ? s - RS : s;
    };
    auto ld3 = [&](int slot, float2 (&x)[3]) {
#pragma unroll
        for (int c = 0; c < 3; ++c) x[c] = ring[(slot * 3 + c) * CVVDP_TSR_THREADS];
    };
    // raw samples of frames it and it + 1, loaded ahead of their use
    unsigned q0a[3] = {0u, 0u, 0u}, q0b[3] = {0u, 0u, 0u}, q1a[3] = {0u, 0u, 0u}, q1b[3] = {0u, 0u, 0u};
    load_raw(0, q0a, q0b);
    load_raw(1, q1a, q1b);
    int rs = 0;  // ring slot of frame `it`
    for (int it = 0; it < NI;) {
        if (it >= FL - 1 && it + 1 < NI) {
            // ---- two output frames t, t+1 per pass over the ring: every slot is read once for both ----
            const int rs1 = slot_add(rs, 1);
            unsigned ca[3] = {q0a[0], q0a[1], q0a[2]}, cb[3] = {q0b[0], q0b[1], q0b[2]};
            unsigned ea[3] = {q1a[0], q1a[1], q1a[2]}, eb[3] = {q1b[0], q1b[1], q1b[2]};
            load_raw(it + 2, q0a, q0b);
            load_raw(it + 3, q1a, q1b);
            convert(it, rs, ca, cb);
            convert(it + 1, rs1, ea, eb);
            // frame t-(FL-1)+k is A_k, frame t-k is B_k: out[t] pairs (A_k, B_k), out[t+1] pairs (A_{k+1}, B_{k-1})
            int sa = slot_add(rs, 2), sb = rs;
            float2 A[3], Bp[3];
            ld3(sa, A);      // A_0
            ld3(rs1, Bp);    // B_{-1} = frame t+1
            float2 o0 = make_float2(0.f, 0.f), o1 = o0, o2 = o0, o3 = o0, u0 = o0, u1 = o0, u2 = o0, u3 = o0;
#pragma unroll 2
            for (int k = 0; k < HP; ++k) {
                const float4 tp = s_taps[k];
                sa = slot_add(sa, 1);
                float2 An[3], Bc[3];
                ld3(sa, An);  // A_{k+1}
                ld3(sb, Bc);  // B_k
                const float2 p0 = add2(A[0], Bc[0]), p1 = add2(A[1], Bc[1]), p2 = add2(A[2], Bc[2]);
                const float2 r0 = add2(An[0], Bp[0]), r1 = add2(An[1], Bp[1]), r2 = add2(An[2], Bp[2]);
                o0 = fma2(bc2(tp.x), p0, o0);
                o1 = fma2(bc2(tp.y), p1, o1);
                o2 = fma2(bc2(tp.z), p2, o2);
                o3 = fma2(bc2(tp.w), p0, o3);
                u0 = fma2(bc2(tp.x), r0, u0);
                u1 = fma2(bc2(tp.y), r1, u1);
                u2 = fma2(bc2(tp.z), r2, u2);
                u3 = fma2(bc2(tp.w), r0, u3);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    A[c] = An[c];
                    Bp[c] = Bc[c];
                }
                sb = sb == 0 ? RS - 1 : sb - 1;
            }
            {   // centre taps: A_HP for out[t], B_{HP-1} (= A_{HP+1}) for out[t+1]
                const float4 tp = s_taps[HP];
                o0 = fma2(bc2(tp.x), A[0], o0);
                o1 = fma2(bc2(tp.y), A[1], o1);
                o2 = fma2(bc2(tp.z), A[2], o2);
                o3 = fma2(bc2(tp.w), A[0], o3);
                u0 = fma2(bc2(tp.x), Bp[0], u0);
                u1 = fma2(bc2(tp.y), Bp[1], u1);
                u2 = fma2(bc2(tp.z), Bp[2], u2);
                u3 = fma2(bc2(tp.w), Bp[0], u3);
            }
            outp[0] = make_float4(o0.x, o1.x, o2.x, o3.x);
            outp[32] = make_float4(o0.y, o1.y, o2.y, o3.y);
            outp += ostep;
            outp[0] = make_float4(u0.x, u1.x, u2.x, u3.x);
            outp[32] = make_float4(u0.y, u1.y, u2.y, u3.y);
            outp += ostep;
            it += 2;
            rs = slot_add(rs, 2);
        } else {
            // ---- one frame: the warm-up frames before the first output, and an odd last output ----
            unsigned ca[3] = {q0a[0], q0a[1], q0a[2]}, cb[3] = {q0b[0], q0b[1], q0b[2]};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                q0a[c] = q1a[c];
                q0b[c] = q1b[c];
            }
            load_raw(it + 2, q1a, q1b);
            convert(it, rs, ca, cb);
            if (it >= FL - 1) {  // uniform: the ring holds frames t-(FL-1) .. t, the oldest two slots ahead of rs
                float2 o0 = make_float2(0.f, 0.f), o1 = o0, o2 = o0, o3 = o0;
                int sa = slot_add(rs, 2), sb = rs;
#pragma unroll 2
                for (int k = 0; k < HP; ++k) {  // tap k <-> frame t-(FL-1)+k, mirrored by frame t-k
                    const float4 tp = s_taps[k];
                    float2 x[3], y[3];
                    ld3(sa, x);
                    ld3(sb, y);
                    const float2 p0 = add2(x[0], y[0]), p1 = add2(x[1], y[1]), p2 = add2(x[2], y[2]);
                    o0 = fma2(bc2(tp.x), p0, o0);
                    o1 = fma2(bc2(tp.y), p1, o1);
                    o2 = fma2(bc2(tp.z), p2, o2);
                    o3 = fma2(bc2(tp.w), p0, o3);
                    sa = slot_add(sa, 1);
                    sb = sb == 0 ? RS - 1 : sb - 1;
                }
                {
                    const float4 tp = s_taps[HP];  // centre tap: sa has arrived at the middle frame
                    float2 x[3];
                    ld3(sa, x);
                    o0 = fma2(bc2(tp.x), x[0], o0);
                    o1 = fma2(bc2(tp.y), x[1], o1);
                    o2 = fma2(bc2(tp.z), x[2], o2);
                    o3 = fma2(bc2(tp.w), x[0], o3);
                }
                outp[0] = make_float4(o0.x, o1.x, o2.x, o3.x);
                outp[32] = make_float4(o0.y, o1.y, o2.y, o3.y);
                outp += ostep;
            }
            it += 1;
            rs = slot_add(rs, 1);
        }
    }
    if (!USE_LUT) publish_input_bits(vbits, a.flags);
    if (it_zero >= 0 && it_zero < NI) {  // uniform per CTA
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) msum += __shfl_xor_sync(0xffffffffu, msum, o);
        if (lane == 0) atomicAdd(a.mean0 + (blockIdx.x & (CVVDP_MEAN_SLOTS - 1)), msum);
    }
}

// =================================================================================================
// Gaussian pyramid reduce  (lpyr_dec.py:186-211): zero-padded 5-tap stride-2 passes (rows, then
// columns) with the reference's edge fix-ups, including the parity quirk at line 206 (the ROW count
// selects the right-edge rule of the column pass).
// =================================================================================================
struct ReduceArgs {
    const float4 *in;
    float4 *out;
    int h, w, hc, wc;
};
#define CVVDP_RTX 32
#define CVVDP_RTY 8
#define CVVDP_RIW (2 * CVVDP_RTX + 3)
#define CVVDP_RIH (2 * CVVDP_RTY + 3)

__global__ void __launch_bounds__(256) k_reduce(const ReduceArgs a) {
    __shared__ float4 s_in[CVVDP_RIH][CVVDP_RIW + 1];
    __shared__ float4 s_ya[CVVDP_RTY][CVVDP_RIW + 1];
    const float K0 = 0.05f, K1 = 0.25f, K2 = 0.4f;  // [K0 K1 K2 K1 K0], lpyr_dec.py:179
    const int tid = threadIdx.x;
    const int plane = blockIdx.z;
    const int ox0 = blockIdx.x * CVVDP_RTX, oy0 = blockIdx.y * CVVDP_RTY;
    const int ix0 = 2 * ox0 - 2, iy0 = 2 * oy0 - 2;
    const float4 *src = a.in + (long long)plane * a.h * a.w;
    for (int i = tid; i < CVVDP_RIH * CVVDP_RIW; i += 256) {
        const int r = i / CVVDP_RIW, c = i - r * CVVDP_RIW;
        const int gy = iy0 + r, gx = ix0 + c;
        float4 v = f4(0.f);
        if (gy >= 0 && gy < a.h && gx >= 0 && gx < a.w) v = src[(long long)gy * a.w + gx];
        s_in[r][c] = v;
    }
    __syncthreads();
    const bool rows_odd = (a.h & 1) != 0;
    for (int i = tid; i < CVVDP_RTY * CVVDP_RIW; i += 256) {
        const int oy = i / CVVDP_RIW, c = i - oy * CVVDP_RIW;
        const int goy = oy0 + oy;
        float4 acc = f4(0.f);
        if (goy < a.hc) {
            const int r = 2 * oy;
            acc = K0 * s_in[r][c];
            acc = fma4(K1, s_in[r + 1][c], acc);
            acc = fma4(K2, s_in[r + 2][c], acc);
            acc = fma4(K1, s_in[r + 3][c], acc);
            acc = fma4(K0, s_in[r + 4][c], acc);
            if (goy == 0) {  // l.195: x~(-1) = x(0), x~(-2) = x(1)
                acc = fma4(K1, s_in[0 - iy0][c], acc);
                acc = fma4(K0, s_in[min(1, a.h - 1) - iy0][c], acc);
            }
            if (goy == a.hc - 1) {  // l.196-199
                if (rows_odd) {
                    acc = fma4(K1, s_in[a.h - 1 - iy0][c], acc);
                    acc = fma4(K0, s_in[max(a.h - 2, 0) - iy0][c], acc);
                } else {
                    acc = fma4(K0, s_in[a.h - 1 - iy0][c], acc);
                }
            }
        }
        s_ya[oy][c] = acc;
    }
    __syncthreads();
    const int ox = tid % CVVDP_RTX, oy = tid / CVVDP_RTX;
    const int gox = ox0 + ox, goy = oy0 + oy;
    if (gox < a.wc && goy < a.hc) {
        const int c = 2 * ox;
        float4 acc = K0 * s_ya[oy][c];
        acc = fma4(K1, s_ya[oy][c + 1], acc);
        acc = fma4(K2, s_ya[oy][c + 2], acc);
        acc = fma4(K1, s_ya[oy][c + 3], acc);
        acc = fma4(K0, s_ya[oy][c + 4], acc);
        if (gox == 0) {  // l.205
            acc = fma4(K1, s_ya[oy][0 - ix0], acc);
            acc = fma4(K0, s_ya[oy][min(1, a.w - 1) - ix0], acc);
        }
        if (gox == a.wc - 1) {  // l.206-209: the parity of the ROW count chooses the rule
            if (rows_odd) {
                acc = fma4(K1, s_ya[oy][a.w - 1 - ix0], acc);
                acc = fma4(K0, s_ya[oy][max(a.w - 2, 0) - ix0], acc);
            } else {
                acc = fma4(K0, s_ya[oy][a.w - 1 - ix0], acc);
            }
        }
        a.out[(long long)plane * a.hc * a.wc + (long long)goy * a.wc + gox] = acc;
    }
}

// Persistent TMA version of the reduce: each CTA owns a CONTIGUOUS range of output tiles of 30 x 8 pixels,
// enumerated column-major inside a plane (ty fastest), so that consecutive tiles of a CTA are vertical neighbours:
// the three halo rows a tile shares with its predecessor were fetched by the same SM microseconds earlier and
// come from the L2 instead of DRAM (round 1 walked the tiles row-major with a grid stride and re-read every halo
// row from DRAM: 1.19x the algorithmic bytes).  The 63 x 19 input box of the NEXT tile is fetched by one
// cp.async.bulk.tensor while the current one is filtered (two shared-memory buffers, one mbarrier each).  TMA's
// zero fill outside the image is exactly the zero padding of the reference's strided conv2d; the edge fix-ups
// are the same as in k_reduce.  The vertical pass stores even and odd columns apart, so that the stride-2
// column pass reads consecutive 128-bit words (no bank conflicts).
#define CVVDP_R2_TX 30
#define CVVDP_R2_IW (2 * CVVDP_R2_TX + 3)  // 63 pixels = 252 floats (TMA box limit 256)
#define CVVDP_R2_EP 36                     // float4 pitch of the even-column half of a ya row (32 + 4: the odd half
                                           // then starts 64 bytes off a 128-byte boundary)
struct Reduce2Args {
    TensorMap3D tm_in;  // fp32 view [planes][h][4w], box {252, 2 TY + 3, 1}
    float4 *out;
    int h, w, hc, wc, planes;
};
template <int TY>
struct Reduce2Smem {
    static constexpr int IH = 2 * TY + 3;                                    // 19 input rows
    static constexpr int BUF = (IH * CVVDP_R2_IW * 16 + 127) / 128 * 128 / 16;  // float4 per buffer, 128-byte multiple
    float4 in[2][BUF];
    float4 ya[TY][CVVDP_R2_EP + 32];  // [oy][0..31]: columns 0,2,4,...; [oy][36..67]: columns 1,3,5,...
    unsigned long long bar[2];
};
template <int TY>
__global__ void __launch_bounds__(256) k_reduce2(const __grid_constant__ Reduce2Args a) {
    CVVDP_DYN_SMEM(smem_raw);
    Reduce2Smem<TY> &sm = *reinterpret_cast<Reduce2Smem<TY> *>(smem_raw);
    constexpr int IH = Reduce2Smem<TY>::IH;
    const float K0 = 0.05f, K1 = 0.25f, K2 = 0.4f;
    const int tid = threadIdx.x;
    const int ntx = (a.wc + CVVDP_R2_TX - 1) / CVVDP_R2_TX, nty = (a.hc + TY - 1) / TY;
    const long long total = (long long)ntx * nty * a.planes;
    // this CTA's contiguous tile range; tile index t = (plane * ntx + tx) * nty + ty
    const long long t_begin = total * blockIdx.x / gridDim.x, t_end = total * (blockIdx.x + 1) / gridDim.x;
    const bool rows_odd = (a.h & 1) != 0;
    if (tid == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
    }
    __syncthreads();
    auto issue = [&](long long t, int buf) {
        const int plane = (int)(t / (ntx * nty)), rem = (int)(t - (long long)plane * (ntx * nty));
        const int tx = rem / nty, ty = rem - tx * nty;
        fence_proxy_async();
        mbar_expect_tx(&sm.bar[buf], IH * CVVDP_R2_IW * 16);
        tma_load_3d(&sm.in[buf][0], &a.tm_in, 4 * (2 * tx * CVVDP_R2_TX - 2), 2 * ty * TY - 2, plane, &sm.bar[buf]);
        mbar_emu_complete(&sm.bar[buf]);
    };
    if (tid == 0 && t_begin < t_end) issue(t_begin, 0);
    int k = 0;
    for (long long t = t_begin; t < t_end; ++t, ++k) {
        const int buf = k & 1;
        if (tid == 0 && t + 1 < t_end) issue(t + 1, buf ^ 1);
        mbar_wait(&sm.bar[buf], (unsigned)((k >> 1) & 1));
        const float4 *in = sm.in[buf];
        const int plane = (int)(t / (ntx * nty)), rem = (int)(t - (long long)plane * (ntx * nty));
        const int tx = rem / nty, ty = rem - tx * nty;
        const int ox0 = tx * CVVDP_R2_TX, oy0 = ty * TY;
        const int ix0 = 2 * ox0 - 2, iy0 = 2 * oy0 - 2;
        for (int i = tid; i < TY * CVVDP_R2_IW; i += 256) {
            const int oy = i / CVVDP_R2_IW, c = i - oy * CVVDP_R2_IW;
            const int goy = oy0 + oy;
            float4 acc = f4(0.f);
            if (goy < a.hc) {
                const float4 *col = in + (2 * oy) * CVVDP_R2_IW + c;
                acc = K0 * col[0];
                acc = fma4(K1, col[CVVDP_R2_IW], acc);
                acc = fma4(K2, col[2 * CVVDP_R2_IW], acc);
                acc = fma4(K1, col[3 * CVVDP_R2_IW], acc);
                acc = fma4(K0, col[4 * CVVDP_R2_IW], acc);
                if (goy == 0) {  // lpyr_dec.py:195
                    acc = fma4(K1, in[(0 - iy0) * CVVDP_R2_IW + c], acc);
                    acc = fma4(K0, in[(min(1, a.h - 1) - iy0) * CVVDP_R2_IW + c], acc);
                }
                if (goy == a.hc - 1) {  // lpyr_dec.py:196-199
                    if (rows_odd) {
                        acc = fma4(K1, in[(a.h - 1 - iy0) * CVVDP_R2_IW + c], acc);
                        acc = fma4(K0, in[(max(a.h - 2, 0) - iy0) * CVVDP_R2_IW + c], acc);
                    } else {
                        acc = fma4(K0, in[(a.h - 1 - iy0) * CVVDP_R2_IW + c], acc);
                    }
                }
            }
            sm.ya[oy][(c & 1) * CVVDP_R2_EP + (c >> 1)] = acc;
        }
        __syncthreads();
        // column c of the vertical pass lives at ya[oy][(c & 1) * EP + (c >> 1)]
        auto ya_at = [&](int oy, int c) -> const float4 & { return sm.ya[oy][(c & 1) * CVVDP_R2_EP + (c >> 1)]; };
        for (int o = tid; o < CVVDP_R2_TX * TY; o += 256) {  // one pass for TY = 8 (240 outputs)
            const int ox = o % CVVDP_R2_TX, oy = o / CVVDP_R2_TX;
            const int gox = ox0 + ox, goy = oy0 + oy;
            if (gox < a.wc && goy < a.hc) {
                const float4 *ev = &sm.ya[oy][ox], *od = &sm.ya[oy][CVVDP_R2_EP + ox];  // columns 2 ox + {0,2,4} / {1,3}
                float4 acc = K0 * ev[0];
                acc = fma4(K1, od[0], acc);
                acc = fma4(K2, ev[1], acc);
                acc = fma4(K1, od[1], acc);
                acc = fma4(K0, ev[2], acc);
                if (gox == 0) {  // lpyr_dec.py:205
                    acc = fma4(K1, ya_at(oy, 0 - ix0), acc);
                    acc = fma4(K0, ya_at(oy, min(1, a.w - 1) - ix0), acc);
                }
                if (gox == a.wc - 1) {  // lpyr_dec.py:206-209: the parity of the ROW count chooses the rule
                    if (rows_odd) {
                        acc = fma4(K1, ya_at(oy, a.w - 1 - ix0), acc);
                        acc = fma4(K0, ya_at(oy, max(a.w - 2, 0) - ix0), acc);
                    } else {
                        acc = fma4(K0, ya_at(oy, a.w - 1 - ix0), acc);
                    }
                }
                a.out[(long long)plane * a.hc * a.wc + (long long)goy * a.wc + gox] = acc;
            }
        }
        __syncthreads();  // ya and in[buf] are free again
    }
}

// =================================================================================================
// Fused band kernel (one launch per pyramid level i < L-1):
//   expand(g_{i+1}) -> Laplacian -> Weber contrast (lpyr_dec.py:386-408) -> castleCSF LUT
//   (csf.py:28-51) -> mult-mutual masking with the 13x13 phase-uncertainty Gaussian, cross-channel
//   pooling and soft clamp (cvvdp_metric.py:817-856, 963-971, 753-764, 945-950) -> spatial
//   p-norm partial sums (cvvdp_metric.py:722, 1032-1048) [-> per-band heat-map plane, 724-734].
// The per-pixel helpers below are used by the strip-marching kernel k_band2 (the first
// version of this file had a 32x32-tile kernel on top of them; it was 2.7x slower and is gone).
// =================================================================================================
#define CVVDP_BHALO 6

struct BandArgs {
    const float4 *fine;    // level i   [pairs*2][h*w]
    const float4 *coarse;  // level i+1 [pairs*2][hc*wc]
    const float4 *lut;     // [32] per-level CSF rows of the 4 channels, pre-scaled: row*log2(10)+log2(sens*gain)
    float *partials;       // [pairs][tiles][4]
    float *hm;             // [pairs][h*w] per-band heat-map plane or null
    float4 *feat;          // feature mode: [3][pairs][h*w] planes |T|S, |R|S, D, or null
    long long feat_plane;  // float4 elements between two of those three plane sets (pairs * h * w)
    float inv_gain[4];     // 1 / masking gain: the CSF rows carry the gain, the features do not (cvvdp_ml_metric.py:352)
    int h, w, hc, wc;
    int do_blur;
    float mul;             // get_band: 1 for band 0, 2 for middle bands (lpyr_dec.py:60-66)
    float lut_a, lut_b;    // LUT index = clamp(log2(L_bkg) * lut_a + lut_b, 0, 31)
    // packed-operand constants: 16-byte aligned so that pairs load straight into aligned uniform-register pairs
    alignas(16) float X[16];  // 2^xcm_weights [source][masked]
    alignas(16) float q[4];
    alignas(16) float kern[2 * CVVDP_BHALO + 1];
    float mc;              // 10^mask_c
    float p;
    float dmax;            // 10^d_max
    float inv_dmax;        // 1 / dmax
    float eps;
    float beta;
    float hm_w[4];         // heat map: channel weights (x image_int)
    float hm_beta, hm_scale;  // beta_tch, 1/band_mul (lpyr_dec.py:308-314)
    int seg_rows;          // k_band2: rows per vertical segment (multiple of 8)
    int use_tma;           // k_band2: stage the rows with TMA (else cp.async)
    TensorMap3D tm_fine;   // fp32 view [planes][h][4w] of level i,   box {4 EW, 8, 2}
    TensorMap3D tm_coarse; // fp32 view [planes][hc][4wc] of level i+1, box {4 CC, 6, 2}
    // fused variant (k_band2f): level i+1 is COMPUTED here and written to coarse_out
    float4 *coarse_out;    // level i+1 [pairs*2][hc*wc]
    TensorMap3D tm_fine_a; // fp32 view of level i, box {256, 11, 2}: fine columns x0-10 .. x0+53
    TensorMap3D tm_fine_b; // same view, box {16, 11, 2}: fine columns x0+54 .. x0+57
};

__device__ __forceinline__ int reflect_idx(int i, int n) {  // torch 'reflect' padding
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

// Contrast / CSF / mutual-masking inputs of one pixel (phase 1).
template <bool FEAT = false>
__device__ __forceinline__ void band_pixel(const BandArgs &a, const float4 *s_lut, float4 gt, float4 gr, float4 et,
                                           float4 er, float4 &mm, float4 &df, float4 *feat_t = nullptr, float4 *feat_r = nullptr) {
    const float4 lt = gt - et, lr = gr - er;  // Laplacian (lpyr_dec.py:387)
    const float Lt = fmaxf(et.x, 0.01f), Lr = fmaxf(er.x, 0.01f);  // l.394
    const float it = a.mul * f_rcp(Lt), ir = a.mul * f_rcp(Lr);
    const float cl = 1000.0f * a.mul;  // clamp(max=1000) before the band multiplier
    const float4 ctu = it * lt, cru = ir * lr;  // packed multiplies
    const float4 ct = make_float4(fminf(ctu.x, cl), fminf(ctu.y, cl), fminf(ctu.z, cl), fminf(ctu.w, cl));
    const float4 cr = make_float4(fminf(cru.x, cl), fminf(cru.y, cl), fminf(cru.z, cl), fminf(cru.w, cl));
    // CSF: always from the reference background (cvvdp_metric.py:709); interp.py:55-60, 92-100
    float ind = fminf(fmaxf(fmaf(f_lg2(Lr), a.lut_a, a.lut_b), 0.f), (float)(CVVDP_CSF_LUT_N - 1));
#ifdef __CUDA_ARCH__
    // floor through the FP32 adder (round-toward-zero add of 2^23) instead of F2I + I2F: the conversion
    // instructions share the quarter-rate XU pipe with the MUFU operations this kernel is short of
    const float tf = __fadd_rz(ind, 8388608.0f);
    const int i0 = __float_as_int(tf) & (CVVDP_CSF_LUT_N - 1);
    const float fr = ind - (tf - 8388608.0f);
#else
    const int i0 = (int)ind;
    const float fr = ind - (float)i0;
#endif
    const int i1 = min(i0 + 1, CVVDP_CSF_LUT_N - 1);
    const float4 va = s_lut[i0], vb = s_lut[i1];
    const float w0 = 1.f - fr;
    const float4 lv = fma4(fr, vb, w0 * va);  // v[i0] (1 - fr) + v[i1] fr, packed
    const float4 S = make_float4(f_ex2(lv.x), f_ex2(lv.y), f_ex2(lv.z), f_ex2(lv.w));
    const float2 S01 = make_float2(S.x, S.y), S23 = make_float2(S.z, S.w);
    const float2 T01 = mul2(make_float2(ct.x, ct.y), S01), T23 = mul2(make_float2(ct.z, ct.w), S23);
    const float2 R01 = mul2(make_float2(cr.x, cr.y), S01), R23 = mul2(make_float2(cr.z, cr.w), S23);
    mm = make_float4(fminf(fabsf(T01.x), fabsf(R01.x)), fminf(fabsf(T01.y), fabsf(R01.y)),
                     fminf(fabsf(T23.x), fabsf(R23.x)), fminf(fabsf(T23.y), fabsf(R23.y)));
    const float2 d01 = add2(T01, make_float2(-R01.x, -R01.y)), d23 = add2(T23, make_float2(-R23.x, -R23.y));
    df = make_float4(fabsf(d01.x), fabsf(d01.y), fabsf(d23.x), fabsf(d23.y));
    if (FEAT) {  // |T_f| S and |R_f| S without the masking gain
        *feat_t = make_float4(fabsf(T01.x) * a.inv_gain[0], fabsf(T01.y) * a.inv_gain[1], fabsf(T23.x) * a.inv_gain[2], fabsf(T23.y) * a.inv_gain[3]);
        *feat_r = make_float4(fabsf(R01.x) * a.inv_gain[0], fabsf(R01.y) * a.inv_gain[1], fabsf(R23.x) * a.inv_gain[2], fabsf(R23.y) * a.inv_gain[3]);
    }
}

__device__ __forceinline__ float spow_fast(float x, float p, float eps, float eps_p) {
    return f_pow(x + eps, p) - eps_p;  // safe_pow, cvvdp_metric.py:77-84
}

// Masking + clamp + pooling term of one pixel (phase 4).  m = blurred mutual-masking signal.
__device__ __forceinline__ float4 band_mask(const BandArgs &a, float4 m, float4 df, const float *eps_q, float eps_p) {
    // channel pairs (A-sust, RG) and (YV, A-trans) go through packed fp32x2 arithmetic; MUFU stays scalar
    const float2 mc2 = bc2(a.mc), eps2 = bc2(a.eps);
    const float2 b01 = fma2(make_float2(m.x, m.y), mc2, eps2), b23 = fma2(make_float2(m.z, m.w), mc2, eps2);
    const float2 e01 = mul2(make_float2(a.q[0], a.q[1]), make_float2(f_lg2(b01.x), f_lg2(b01.y)));
    const float2 e23 = mul2(make_float2(a.q[2], a.q[3]), make_float2(f_lg2(b23.x), f_lg2(b23.y)));
    const float2 t01 = add2(make_float2(f_ex2(e01.x), f_ex2(e01.y)), make_float2(-eps_q[0], -eps_q[1]));
    const float2 t23 = add2(make_float2(f_ex2(e23.x), f_ex2(e23.y)), make_float2(-eps_q[2], -eps_q[3]));
    // 1 + M[c], M[c] = sum_i t_i X[i][c]  (cvvdp_metric.py:758-760), two masked channels per operation
    const float2 one2 = bc2(1.f);
    float2 M01 = fma2(bc2(t01.x), make_float2(a.X[0], a.X[1]), one2);
    float2 M23 = fma2(bc2(t01.x), make_float2(a.X[2], a.X[3]), one2);
    M01 = fma2(bc2(t01.y), make_float2(a.X[4], a.X[5]), M01);
    M23 = fma2(bc2(t01.y), make_float2(a.X[6], a.X[7]), M23);
    M01 = fma2(bc2(t23.x), make_float2(a.X[8], a.X[9]), M01);
    M23 = fma2(bc2(t23.x), make_float2(a.X[10], a.X[11]), M23);
    M01 = fma2(bc2(t23.y), make_float2(a.X[12], a.X[13]), M01);
    M23 = fma2(bc2(t23.y), make_float2(a.X[14], a.X[15]), M23);
    const float2 d01 = add2(make_float2(df.x, df.y), eps2), d23 = add2(make_float2(df.z, df.w), eps2);
    const float2 p2 = bc2(a.p), nep = bc2(-eps_p);
    const float2 g01 = mul2(p2, make_float2(f_lg2(d01.x), f_lg2(d01.y))), g23 = mul2(p2, make_float2(f_lg2(d23.x), f_lg2(d23.y)));
    const float2 P01 = add2(make_float2(f_ex2(g01.x), f_ex2(g01.y)), nep), P23 = add2(make_float2(f_ex2(g23.x), f_ex2(g23.y)), nep);
    // D_u = P / (1 + M);  D = Dmax * D_u / (Dmax + D_u)  ==  P / ((1 + M) + P / Dmax)   (cvvdp_metric.py:855, 948-950)
    const float2 idm2 = bc2(a.inv_dmax);
    const float2 n01 = fma2(P01, idm2, M01), n23 = fma2(P23, idm2, M23);
    const float2 D01 = mul2(P01, make_float2(f_rcp(n01.x), f_rcp(n01.y)));
    const float2 D23 = mul2(P23, make_float2(f_rcp(n23.x), f_rcp(n23.y)));
    return make_float4(D01.x, D01.y, D23.x, D23.y);
}

// =================================================================================================
// Fused band kernel, strip-marching.
// A CTA of 128 threads owns a vertical strip of SW = EW - 12 columns (EW with the +-6 halo of the 13x13
// Gaussian) of one (item, frame) and marches down a segment of rows, 8 rows per step, keeping rolling
// windows in shared memory:
//   fine, crs : the 8 fine rows of this step and the 6 coarse rows under them (TMA stage, one step ahead)
//   mm        : min(|T'|,|R'|) of the 8 rows of this step (EW columns)
//   hb        : ring of the last 32 rows of the horizontally blurred mm (SW columns)
//   df        : ring of the last 16 rows of |T'-R'| waiting for their blurred mask
// so the 13x13 Gaussian costs 13+13 taps and the vertical halo is paid once per segment.
// Step k: rows A = [a0, a0+8) get contrast/CSF/mm/df (2x2 quads, one per thread) and their horizontal
// blur; rows C = [a0-6, a0+2) -- whose 13-row window is now complete -- get the vertical blur,
// masking, clamp and pooling.
// Geometry (template EW): 64 -> 52-column strips, phases B/C use 104 of 128 threads; 60 -> 48-column
// strips, phase A uses 120 threads and phases B/C exactly three warps (no idle lanes in a running warp).
// Shipped configuration: EW = 60, CF and LAG on (round-2 A/B on one box, 4K x 120 frames, level 0: 17.96 ms for
// round 1's kernel, 17.52 with EW = 60 + CF, 17.11 with LAG added; profiles/r02_ab_band_variants_pa_lag.txt).
// The kernel is bound by the FP32 pipe with additive dispatch interference from MUFU / ALU-pipe instructions
// and a ~70 % busy shared-memory pipe; profiles/r02_microbench_smsp_pipes.txt has the measured per-instruction
// costs (FFMA2 2 cycles, ALU-pipe op 2, MUFU 8, LDS.128 16 per scheduler) behind DESIGN.md section 5.
// =================================================================================================
#define CVVDP_B2_RB 8
#define CVVDP_B2_THREADS 128
#define CVVDP_B2_HBR 32
#define CVVDP_B2_DFR 16
#define CVVDP_B2_CR (CVVDP_B2_RB / 2 + 2)  // 6 coarse rows per step

template <int EW>
struct B2Geom {
    static constexpr int SW = EW - 2 * CVVDP_BHALO;          // useful columns: 52 / 48
    static constexpr int QW = EW / 2;                        // quads per row: 32 / 30
    static constexpr int CC = EW / 2 + 2;                    // coarse columns: 34 / 32
    static constexpr int A_THREADS = QW * (CVVDP_B2_RB / 2);  // 128 / 120
    static constexpr int B_TASKS = CVVDP_B2_RB * (SW / 4);   // 104 / 96
    static constexpr int C_TASKS = 2 * SW;                   // 104 / 96
    static_assert(EW % 4 == 0 && SW % 4 == 0 && A_THREADS <= CVVDP_B2_THREADS && 4 * EW <= 256, "band strip geometry");
};

template <int EW, int DFR = CVVDP_B2_DFR>
struct Band2Smem {
    float4 lut[CVVDP_CSF_LUT_N];                      // 512 B: keeps the TMA destinations 128-byte aligned
    float4 crs[2][CVVDP_B2_CR][B2Geom<EW>::CC];       // coarse rows of the current step (TMA / cp.async stage)
    float4 fine[2][CVVDP_B2_RB][EW];                  // fine rows of the current step   (TMA / cp.async stage)
    float4 mm[CVVDP_B2_RB][EW + 1];
    float4 hb[CVVDP_B2_HBR][B2Geom<EW>::SW + 1];
    float4 df[DFR][B2Geom<EW>::SW];
    float red[CVVDP_B2_THREADS / 32][4];
    unsigned long long bar;                           // mbarrier of the TMA stage
};

// Asynchronous stage of one step: the 8 fine rows [a0, a0+8) x EW columns of both videos and the 6
// coarse rows under them (replicate-clamped, lpyr_dec.py:136-141).  Issued one step ahead.
template <int EW, int DFR>
__device__ __forceinline__ void band2_stage(const BandArgs &a, Band2Smem<EW, DFR> &sm, const float4 *fine_t, const float4 *crs_g,
                                            long long npix, long long ncpix, int a0, int a_end, int ex0, int tid, int pair) {
    constexpr int CC = B2Geom<EW>::CC;
    const int cy0 = a0 / 2 - 1, cx0 = ex0 / 2 - 1;
    if (a.use_tma) {  // two bulk tensor copies issued by one thread; out-of-range elements arrive as zeros
        if (tid == 0) {
            fence_proxy_async();
            mbar_expect_tx(&sm.bar, (unsigned)(sizeof(float4) * 2 * (CVVDP_B2_CR * CC + CVVDP_B2_RB * EW)));
            tma_load_3d(&sm.fine[0][0][0], &a.tm_fine, 4 * ex0, a0, 2 * pair, &sm.bar);
            tma_load_3d(&sm.crs[0][0][0], &a.tm_coarse, 4 * cx0, cy0, 2 * pair, &sm.bar);
            mbar_emu_complete(&sm.bar);
        }
        return;
    }
    for (int i = tid; i < 2 * CVVDP_B2_CR * CC; i += CVVDP_B2_THREADS) {
        const int v = i / (CVVDP_B2_CR * CC), rem = i - v * (CVVDP_B2_CR * CC);
        const int r = rem / CC, c = rem - r * CC;
        const int cy = min(max(cy0 + r, 0), a.hc - 1), cx = min(max(cx0 + c, 0), a.wc - 1);
        cp_async16(&sm.crs[v][r][c], crs_g + v * ncpix + (long long)cy * a.wc + cx);
    }
    for (int i = tid; i < CVVDP_B2_RB * EW; i += CVVDP_B2_THREADS) {
        const int r = i / EW, c = i - r * EW;
        const int gx = ex0 + c, gy = a0 + r;
        if (gx >= 0 && gx < a.w && gy < a_end) {
            const float4 *src = fine_t + (long long)gy * a.w + gx;
            cp_async16(&sm.fine[0][r][c], src);
            cp_async16(&sm.fine[1][r][c], src + npix);
        }
    }
    cp_async_commit();
}

// Template flags: EW = strip geometry (above); CF = conflict-free phase A (below); BLUR = phase-uncertainty
// Gaussian on (off only for levels with h <= 6 or w <= 6, Q7); HM = write the per-band heat-map plane;
// BETA2 = spatial pooling exponent is exactly 2 (shipped value); FEAT = feature mode: additionally write
// |T|S, |R|S (phase A) and D (phase C) of every pixel of the segment's interior to three planes that
// k_feature_pool turns into the per-patch statistics of the ML heads.
// The per-pixel bodies of phases A and C are straight-line code (no per-pixel branches): rows and columns
// outside the segment are computed on whatever the stage holds and discarded by a select, so the compiler
// can interleave the MUFU chains of a thread's four pixels.
// CF: a thread's quad covers fine columns 2qx and 2qx+1, so the eight lanes of a quarter-warp touch 128-bit
// words at a stride of two -- a two-way bank conflict on every fine-row load and every mm / df store of phase
// A (ncu, round 1: 20 % of all shared-memory wavefronts of the kernel).  With CF the lanes whose bit 2 is set
// take the ODD column of their quad first and the even one second; the eight lanes of a quarter-warp then
// cover eight distinct 16-byte bank groups.  The column parity only enters the horizontal expand weights
// ((.1,.8,.1) even, (0,.5,.5) odd), which become per-lane values: bit-identical results, one extra
// packed multiply per (row, video).
// LAG: phase C trails phase B by 16 rows instead of 6, so every row of its 13-row window was written in an EARLIER
// step and phases B and C need no barrier between them (two barriers per step instead of three; the FP32-only
// horizontal blur of some warps overlaps the MUFU-heavy masking of others).  The |T'-R'| ring then holds 24 rows.
template <bool LAG>
struct B2Lag {
    static constexpr int DFR = LAG ? 24 : CVVDP_B2_DFR;
};
template <int EW, bool CF, bool LAG, bool BLUR, bool HM, bool BETA2, bool FEAT = false>
__global__ void __launch_bounds__(CVVDP_B2_THREADS, 3) k_band2(const __grid_constant__ BandArgs a) {
    CVVDP_DYN_SMEM(smem_raw);
    constexpr bool LAGC = LAG && BLUR;           // without the blur phase C reads the rows of its own step
    constexpr int DFR = B2Lag<LAG>::DFR;
    Band2Smem<EW, DFR> &sm = *reinterpret_cast<Band2Smem<EW, DFR> *>(smem_raw);
    constexpr int SW = B2Geom<EW>::SW, QW = B2Geom<EW>::QW, CC = B2Geom<EW>::CC;
    const int tid = threadIdx.x;
    const int pair = blockIdx.z;
    constexpr int hal = BLUR ? CVVDP_BHALO : 0;
    const int x0 = blockIdx.x * SW, ex0 = x0 - hal;  // even
    const int ys = blockIdx.y * a.seg_rows, ye = min(ys + a.seg_rows, a.h);
    const int y_begin = max(ys - hal, 0);                     // even
    const int a_end = min(ye + hal, a.h);                     // rows [y_begin, a_end) feed this segment
    const long long npix = (long long)a.h * a.w, ncpix = (long long)a.hc * a.wc;
    const float4 *fine_t = a.fine + (long long)pair * 2 * npix;
    const float4 *crs_g = a.coarse + (long long)pair * 2 * ncpix;
    const bool x_edge = (ex0 < 0) || (x0 + SW + hal > a.w);  // strip touches the left/right border

    if (a.use_tma) {
        if (tid == 0) mbar_init(&sm.bar, 1);
        __syncthreads();
    }
    unsigned tma_phase = 0;
    const int cx0 = ex0 / 2 - 1;
    band2_stage<EW, DFR>(a, sm, fine_t, crs_g, npix, ncpix, y_begin, a_end, ex0, tid, pair);
    if (tid < CVVDP_CSF_LUT_N) sm.lut[tid] = a.lut[tid];
    float eps_q[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) eps_q[c] = f_pow(a.eps, a.q[c]);
    const float eps_p = f_pow(a.eps, a.p);
    const float eps_b = BETA2 ? 0.f : f_pow(a.eps, a.beta);
    float4 acc = f4(0.f);
    // loop-invariant thread roles
    const int qy = tid / QW, qx = tid - qy * QW;                               // phase A: one 2x2 quad
    const int b_r = tid % CVVDP_B2_RB, b_xg = tid / CVVDP_B2_RB;               // phase B: row, group of 4 columns
    const int c_ix = tid % SW, c_rg = tid / SW;                                // phase C: column, group of 4 rows
    const int a_gx = ex0 + 2 * qx;
    const bool a_cols = a_gx + 1 >= 0 && a_gx < a.w && tid < B2Geom<EW>::A_THREADS;
    // CF: column order inside the quad and the matching horizontal expand weights of this lane
    const int sw_odd = CF ? ((tid >> 2) & 1) : 0;  // 1: this lane takes the odd column of its quad first
    const float wA0 = sw_odd ? 0.f : 0.1f, wA1 = sw_odd ? 0.5f : 0.8f, wA2 = sw_odd ? 0.5f : 0.1f;  // first column
    const float wB0 = sw_odd ? 0.1f : 0.f, wB1 = sw_odd ? 0.8f : 0.5f, wB2 = sw_odd ? 0.1f : 0.5f;  // second column

    constexpr int clag = LAGC ? 16 : hal;  // phase C of a step handles rows [a0 - clag, a0 - clag + 8)
    int dslot = 0;                         // LAG: ring slot of row a0 = (a0 - y_begin) mod 24
    for (int a0 = y_begin; a0 - clag < ye; a0 += CVVDP_B2_RB) {
        const bool have_a = a0 < a_end;
        if (a.use_tma) {
            if (have_a) {
                mbar_wait(&sm.bar, tma_phase);
                tma_phase ^= 1u;
                // TMA zero-fills outside the coarse image; the reference pads by replication
                // (lpyr_dec.py:136-141): patch those entries from the nearest valid row/column
                const int cy0 = a0 / 2 - 1;
                if (cy0 < 0 || cy0 + CVVDP_B2_CR > a.hc || cx0 < 0 || cx0 + CC > a.wc) {
                    for (int i = tid; i < 2 * CVVDP_B2_CR * CC; i += CVVDP_B2_THREADS) {
                        const int v = i / (CVVDP_B2_CR * CC), rem = i - v * (CVVDP_B2_CR * CC);
                        const int r = rem / CC, c = rem - r * CC;
                        const int rr = min(max(cy0 + r, 0), a.hc - 1) - cy0, cc = min(max(cx0 + c, 0), a.wc - 1) - cx0;
                        if ((rr != r || cc != c) && rr >= 0 && rr < CVVDP_B2_CR && cc >= 0 && cc < CC)
                            sm.crs[v][r][c] = sm.crs[v][rr][cc];
                    }
                }
            }
        } else {
            cp_async_wait_all();
        }
        __syncthreads();  // the stage of this step has landed; previous phase C is complete
        // ---- phase A: one 2x2 quad per thread: expand, contrast, CSF -> mm, df ----
        if (have_a) {
            const int gy = a0 + 2 * qy;
            if (gy < a_end && a_cols) {
                float4 e[2][4];
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    float4 ve[3], vo[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float4 c0 = sm.crs[v][qy][qx + c], c1 = sm.crs[v][qy + 1][qx + c], c2 = sm.crs[v][qy + 2][qx + c];
                        ve[c] = fma4(0.1f, c2, fma4(0.8f, c1, 0.1f * c0));
                        vo[c] = fma4(0.5f, c2, 0.5f * c1);
                    }
                    if (CF) {  // e[v][k]: k & 1 = first / second column of this lane
                        e[v][0] = fma4(wA2, ve[2], fma4(wA1, ve[1], wA0 * ve[0]));
                        e[v][1] = fma4(wB2, ve[2], fma4(wB1, ve[1], wB0 * ve[0]));
                        e[v][2] = fma4(wA2, vo[2], fma4(wA1, vo[1], wA0 * vo[0]));
                        e[v][3] = fma4(wB2, vo[2], fma4(wB1, vo[1], wB0 * vo[0]));
                    } else {
                        e[v][0] = fma4(0.1f, ve[2], fma4(0.8f, ve[1], 0.1f * ve[0]));
                        e[v][1] = fma4(0.5f, ve[2], 0.5f * ve[1]);
                        e[v][2] = fma4(0.1f, vo[2], fma4(0.8f, vo[1], 0.1f * vo[0]));
                        e[v][3] = fma4(0.5f, vo[2], 0.5f * vo[1]);
                    }
                }
                // pixels beyond the image / segment are evaluated on the stage's fill values and never read back
                // (phase B reflects at the borders, phase C selects); their df slot belongs to rows long consumed
                float4 mm[4], df[4], ft[4], fr[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int ry = 2 * qy + (k >> 1), rx = 2 * qx + ((k & 1) ^ sw_odd);
                    band_pixel<FEAT>(a, sm.lut, sm.fine[0][ry][rx], sm.fine[1][ry][rx], e[0][k], e[1][k], mm[k], df[k], &ft[k], &fr[k]);
                }
                const int ix = ex0 + 2 * qx - x0;  // even; the pair (ix, ix+1) is inside or outside the strip together
                const bool in_strip = ix >= 0 && ix < SW;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int ry = 2 * qy + (k >> 1), cofs = (k & 1) ^ sw_odd, rx = 2 * qx + cofs;
                    sm.mm[ry][rx] = mm[k];
                    if (in_strip) sm.df[LAG ? dslot + ry : ((a0 + ry) & (CVVDP_B2_DFR - 1))][ix + cofs] = df[k];
                    if (FEAT) {  // every pixel belongs to the interior of exactly one (strip, segment)
                        const int py = a0 + ry, px = ex0 + rx;
                        if (in_strip && py >= ys && py < ye && px < a.w) {
                            float4 *dst = a.feat + (long long)pair * npix + (long long)py * a.w + px;
                            dst[0] = ft[k];
                            dst[a.feat_plane] = fr[k];
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ---- prefetch the next step's stage while phases B and C run ----
        if (a0 + CVVDP_B2_RB < a_end) band2_stage<EW, DFR>(a, sm, fine_t, crs_g, npix, ncpix, a0 + CVVDP_B2_RB, a_end, ex0, tid, pair);
        // ---- phase B: horizontal pass of the phase-uncertainty Gaussian for the new rows ----
        if (BLUR && have_a && tid < B2Geom<EW>::B_TASKS) {
            const int gy = a0 + b_r, gxb = x0 + b_xg * 4;
            if (gy < a_end && gxb < a.w) {
                float4 win[2 * CVVDP_BHALO + 4];
                if (x_edge) {
#pragma unroll
                    for (int j = 0; j < 2 * CVVDP_BHALO + 4; ++j) {
                        int lx = reflect_idx(gxb + j - CVVDP_BHALO, a.w) - ex0;
                        lx = min(max(lx, 0), EW - 1);  // only for outputs beyond the image (discarded)
                        win[j] = sm.mm[b_r][lx];
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 2 * CVVDP_BHALO + 4; ++j) win[j] = sm.mm[b_r][b_xg * 4 + j];
                }
                float4 *dst = &sm.hb[gy & (CVVDP_B2_HBR - 1)][b_xg * 4];
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    float4 s = f4(0.f);
#pragma unroll
                    for (int k = 0; k < 2 * CVVDP_BHALO + 1; ++k) s = fma4(a.kern[k], win[o + k], s);
                    dst[o] = s;
                }
            }
        }
        if (!LAGC) __syncthreads();
        // ---- phase C: vertical pass, masking, clamp, pooling for the rows whose window is complete ----
        if (tid < B2Geom<EW>::C_TASKS) {
            const int gx = x0 + c_ix;
            const int cyb = a0 - clag + c_rg * 4;  // 4 consecutive rows per thread
            int dbase = dslot + (LAGC ? 8 : 0) + 4 * c_rg;  // LAG: ring slot of row cyb = (cyb - y_begin) mod 24
            dbase -= dbase >= 24 ? 24 : 0;
            if (gx < a.w && cyb + 3 >= ys && cyb < ye) {
                float4 win[2 * CVVDP_BHALO + 4];
                if (BLUR) {
                    const bool y_edge = (cyb - CVVDP_BHALO < 0) || (cyb + 3 + CVVDP_BHALO >= a.h);
                    if (y_edge) {
#pragma unroll
                        for (int j = 0; j < 2 * CVVDP_BHALO + 4; ++j) {
                            int yy = reflect_idx(cyb + j - CVVDP_BHALO, a.h);
                            yy = min(max(yy, 0), a.h - 1);
                            win[j] = sm.hb[yy & (CVVDP_B2_HBR - 1)][c_ix];
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 2 * CVVDP_BHALO + 4; ++j)
                            win[j] = sm.hb[(cyb + j - CVVDP_BHALO) & (CVVDP_B2_HBR - 1)][c_ix];
                    }
                }
                float4 D[4];
#pragma unroll
                for (int o = 0; o < 4; ++o) {  // straight-line: rows outside [ys, ye) are discarded below
                    const int gy = cyb + o;
                    float4 m;
                    if (BLUR) {
                        m = f4(0.f);
#pragma unroll
                        for (int k = 0; k < 2 * CVVDP_BHALO + 1; ++k) m = fma4(a.kern[k], win[o + k], m);
                    } else {
                        m = sm.mm[(gy - a0) & (CVVDP_B2_RB - 1)][gx - ex0];
                    }
                    D[o] = band_mask(a, m, sm.df[LAG ? dbase + o : (gy & (CVVDP_B2_DFR - 1))][c_ix], eps_q, eps_p);
                }
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const int gy = cyb + o;
                    const bool live = gy >= ys && gy < ye;
                    float4 t;
                    if (BETA2) {  // (D+eps)^2 - eps^2 == D (D + 2 eps), exact at D = 0
                        const float e2 = 2.f * a.eps;
                        t = make_float4(D[o].x * (D[o].x + e2), D[o].y * (D[o].y + e2), D[o].z * (D[o].z + e2), D[o].w * (D[o].w + e2));
                    } else {
                        t = make_float4(f_pow(D[o].x + a.eps, a.beta) - eps_b, f_pow(D[o].y + a.eps, a.beta) - eps_b,
                                        f_pow(D[o].z + a.eps, a.beta) - eps_b, f_pow(D[o].w + a.eps, a.beta) - eps_b);
                    }
                    acc.x += live ? t.x : 0.f;
                    acc.y += live ? t.y : 0.f;
                    acc.z += live ? t.z : 0.f;
                    acc.w += live ? t.w : 0.f;
                    if (FEAT && live) a.feat[2 * a.feat_plane + (long long)pair * npix + (long long)gy * a.w + gx] = D[o];
                    if (HM && live) {
                        const float eb = f_pow(a.eps, a.hm_beta);
                        float s = (f_pow(D[o].x * a.hm_w[0] + a.eps, a.hm_beta) - eb) + (f_pow(D[o].y * a.hm_w[1] + a.eps, a.hm_beta) - eb) +
                                  (f_pow(D[o].z * a.hm_w[2] + a.eps, a.hm_beta) - eb) + (f_pow(D[o].w * a.hm_w[3] + a.eps, a.hm_beta) - eb);
                        const float ib = 1.f / a.hm_beta;
                        a.hm[(long long)pair * npix + (long long)gy * a.w + gx] = (f_pow(s + a.eps, ib) - f_pow(a.eps, ib)) * a.hm_scale;
                    }
                }
            }
        }
        dslot = dslot == 16 ? 0 : dslot + 8;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
        acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if ((tid & 31) == 0) {
        sm.red[tid >> 5][0] = acc.x;
        sm.red[tid >> 5][1] = acc.y;
        sm.red[tid >> 5][2] = acc.z;
        sm.red[tid >> 5][3] = acc.w;
    }
    __syncthreads();
    if (tid < 4) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < CVVDP_B2_THREADS / 32; ++w) s += sm.red[w][tid];
        const int tile = blockIdx.y * gridDim.x + blockIdx.x, ntiles = gridDim.x * gridDim.y;
        a.partials[((long long)pair * ntiles + tile) * 4 + tid] = s;
    }
}

// =================================================================================================
// Fused band kernel WITH the Gaussian-pyramid reduce ("one kernel per level", BASELINE north_star):
// level i+1 is computed from the fine rows this kernel stages anyway, used for the expand / Laplacian of level i
// straight from shared memory, and written to HBM once for the next level's kernel.  No reduce launches, and every
// level is read from HBM once instead of twice.
// Same strip-marching structure as k_band2 (EW = 60, conflict-free phase A, no lag), with per step:
//   stage : ONE 11-row x 68-column box of fine rows [a0, a0+11) x [x0-10, x0+58) per video (two TMA boxes, 64 + 4
//           columns wide): the 8 rows of the step plus the 3 rows below that the 5-tap reduce of coarse rows
//           a0/2+1 .. a0/2+4 reaches; zero fill outside the image = the zero padding of the reference's conv2d
//   R1    : vertical 5-tap pass, 4 coarse rows x 67 columns x 2 videos (even / odd columns stored apart)
//   R2    : horizontal pass -> 4 x 32 coarse pixels x 2 videos into an 8-row ring; the pixels this CTA OWNS
//           (coarse columns [24 bx, 24 bx + 24), coarse rows of its row segment) also go to HBM
//   A,B,C : as k_band2; phase A reads the coarse ring with indices clamped to the coarse image (= replicate padding)
// A prologue step (reduce only) at a0 = y_begin - 8 produces the two coarse rows above the first step.
// =================================================================================================
#define CVVDP_BF_EW 60
#define CVVDP_BF_FW 68        // staged fine columns
#define CVVDP_BF_FR 11        // staged fine rows
#define CVVDP_BF_VP 36        // float4 pitch of the even-column half of a vertical-pass row (34 used + 2)
#define CVVDP_BF_HBR 24       // ring of horizontally blurred rows (20 live)
struct BandFSmem {
    float4 lut[CVVDP_CSF_LUT_N];
    float4 fine_a[2][CVVDP_BF_FR][64];                 // TMA box A (columns 0..63 of the staged frame)
    float4 fine_b[2][CVVDP_BF_FR][4];                  // TMA box B (columns 64..67)
    float4 crs[2][8][B2Geom<CVVDP_BF_EW>::CC];         // ring of coarse rows, slot = (cy - cbase) & 7
    union {
        float4 mm[CVVDP_B2_RB][CVVDP_BF_FW + 1];       // phase A -> B
        float4 vp[2][4][2 * CVVDP_BF_VP];              // R1 -> R2: [video][coarse row][even half | odd half]
    };
    float4 hb[CVVDP_BF_HBR][B2Geom<CVVDP_BF_EW>::SW + 1];
    float4 df[CVVDP_B2_DFR][B2Geom<CVVDP_BF_EW>::SW];
    float red[CVVDP_B2_THREADS / 32][4];
    unsigned long long bar;
};
#define CVVDP_BF_STAGE_BYTES ((unsigned)(sizeof(float4) * 2 * CVVDP_BF_FR * CVVDP_BF_FW))

template <bool HM, bool BETA2>
__global__ void __launch_bounds__(CVVDP_B2_THREADS, 3) k_band2f(const __grid_constant__ BandArgs a) {
    CVVDP_DYN_SMEM(smem_raw);
    BandFSmem &sm = *reinterpret_cast<BandFSmem *>(smem_raw);
    constexpr int EW = CVVDP_BF_EW, SW = B2Geom<EW>::SW, QW = B2Geom<EW>::QW, CC = B2Geom<EW>::CC;
    constexpr int hal = CVVDP_BHALO;
    const float K0 = 0.05f, K1 = 0.25f, K2 = 0.4f;
    const int tid = threadIdx.x;
    const int pair = blockIdx.z;
    const int x0 = blockIdx.x * SW, ex0 = x0 - hal, fx0 = x0 - 10;  // phase-A frame / staged frame
    const int ys = blockIdx.y * a.seg_rows, ye = min(ys + a.seg_rows, a.h);
    const int y_begin = max(ys - hal, 0);
    const int a_end = min(ye + hal, a.h);
    const long long npix = (long long)a.h * a.w, ncpix = (long long)a.hc * a.wc;
    const bool x_edge = (ex0 < 0) || (x0 + SW + hal > a.w);
    const int cx0 = ex0 / 2 - 1;              // coarse column of ring column 0 (= 24 bx - 4)
    const int cbase = y_begin / 2 - 4;        // ring slot of coarse row cy: (cy - cbase) & 7
    const int own_cy0 = ys / 2, own_cy1 = (ye == a.h) ? a.hc : ye / 2;  // coarse rows this segment writes
    const bool rows_odd = (a.h & 1) != 0;
    float4 *crs_out = a.coarse_out + (long long)pair * 2 * ncpix;

    if (tid == 0) mbar_init(&sm.bar, 1);
    if (tid < CVVDP_CSF_LUT_N) sm.lut[tid] = a.lut[tid];
    __syncthreads();
    unsigned tma_phase = 0;
    auto stage = [&](int a0) {
        if (tid == 0) {
            fence_proxy_async();
            mbar_expect_tx(&sm.bar, CVVDP_BF_STAGE_BYTES);
            tma_load_3d(&sm.fine_a[0][0][0], &a.tm_fine_a, 4 * fx0, a0, 2 * pair, &sm.bar);
            tma_load_3d(&sm.fine_b[0][0][0], &a.tm_fine_b, 4 * (fx0 + 64), a0, 2 * pair, &sm.bar);
            mbar_emu_complete(&sm.bar);
        }
    };
    stage(y_begin - CVVDP_B2_RB);  // prologue step: reduce only
    float eps_q[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) eps_q[c] = f_pow(a.eps, a.q[c]);
    const float eps_p = f_pow(a.eps, a.p);
    const float eps_b = BETA2 ? 0.f : f_pow(a.eps, a.beta);
    float4 acc = f4(0.f);
    const int qy = tid / QW, qx = tid - qy * QW;
    const int b_r = tid % CVVDP_B2_RB, b_xg = tid / CVVDP_B2_RB;
    const int c_ix = tid % SW, c_rg = tid / SW;
    const int a_gx = ex0 + 2 * qx;
    const bool a_cols = a_gx + 1 >= 0 && a_gx < a.w && tid < B2Geom<EW>::A_THREADS;
    const int sw_odd = (tid >> 2) & 1;
    const float wA0 = sw_odd ? 0.f : 0.1f, wA1 = sw_odd ? 0.5f : 0.8f, wA2 = sw_odd ? 0.5f : 0.1f;
    const float wB0 = sw_odd ? 0.1f : 0.f, wB1 = sw_odd ? 0.8f : 0.5f, wB2 = sw_odd ? 0.1f : 0.5f;
    // ---- reduce, vertical 5-tap pass (lpyr_dec.py:186-199) for the coarse rows a0/2+1 .. a0/2+4 of the step whose
    // fine rows [a0, a0+11) are in the stage: one (video, column) task per thread, 2 x 64 columns of box A, then the
    // 2 x 3 columns of box B on six threads.  EDGE: the step holds coarse row 0 or hc-1 (top / bottom fix-ups);
    // everywhere else the pass is straight-line code.  (Running it one step ahead, in the barrier interval of
    // phase C of the previous step, was tried: 25.4 instead of 23.9 ms at level 0.) ----
    auto reduce_column = [&](const float4 *colp, const int pitch, int v, int f, int a0, int cyn, bool edge) {
        float4 x[CVVDP_BF_FR];
#pragma unroll
        for (int r = 0; r < CVVDP_BF_FR; ++r) x[r] = colp[r * pitch];
        float4 *dst = &sm.vp[v][0][(f & 1) * CVVDP_BF_VP + (f >> 1)];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 s = K0 * x[2 * i];
            s = fma4(K1, x[2 * i + 1], s);
            s = fma4(K2, x[2 * i + 2], s);
            s = fma4(K1, x[2 * i + 3], s);
            s = fma4(K0, x[2 * i + 4], s);
            if (edge) {
                const int cy = cyn + i;
                if (cy == 0) {  // l.195: x~(-1) = x(0), x~(-2) = x(1); stage row of image row y is y - a0
                    s = fma4(K1, colp[(0 - a0) * pitch], s);
                    s = fma4(K0, colp[(min(1, a.h - 1) - a0) * pitch], s);
                }
                if (cy == a.hc - 1) {  // l.196-199
                    if (rows_odd) {
                        s = fma4(K1, colp[(a.h - 1 - a0) * pitch], s);
                        s = fma4(K0, colp[(max(a.h - 2, 0) - a0) * pitch], s);
                    } else {
                        s = fma4(K0, colp[(a.h - 1 - a0) * pitch], s);
                    }
                }
            }
            dst[i * 2 * CVVDP_BF_VP] = s;
        }
    };
    auto reduce_rows = [&](int a0) {
        const int cyn = a0 / 2 + 1;  // a0 is even (negative in the prologue: exact division)
        const bool edge = cyn <= 0 || cyn + 3 >= a.hc - 1;  // uniform
        {
            const int v = tid >> 6, f = tid & 63;
            if (edge) reduce_column(&sm.fine_a[v][0][f], 64, v, f, a0, cyn, true);
            else reduce_column(&sm.fine_a[v][0][f], 64, v, f, a0, cyn, false);
        }
        if (tid < 6) {  // staged columns 64..66
            const int v = tid >= 3 ? 1 : 0, f = 64 + tid - 3 * v;
            reduce_column(&sm.fine_b[v][0][f - 64], 4, v, f, a0, cyn, true);
        }
    };
    int hslot0 = 0;  // hb ring slot of row a0 = (a0 - y_begin) mod 24, advanced per step

    for (int a0 = y_begin - CVVDP_B2_RB; a0 - hal < ye; a0 += CVVDP_B2_RB) {
        const bool have_a = a0 < a_end;             // fine rows of this step exist
        const bool do_a = have_a && a0 >= y_begin;  // not the prologue
        if (have_a) {
            mbar_wait(&sm.bar, tma_phase);
            tma_phase ^= 1u;
        }
        __syncthreads();  // the stage has landed; previous phase C is complete (mm / vp are free)
        if (have_a) reduce_rows(a0);
        __syncthreads();
        if (have_a) {
            const int cyn = a0 / 2 + 1;
            // ---- R2: horizontal pass -> coarse ring (+ HBM for the pixels this CTA owns) (lpyr_dec.py:201-209) ----
            // 2 videos x 4 rows x 32 columns = two tasks per thread; EDGE (uniform): the ring holds coarse column 0 or wc-1
            const bool edge_x = cx0 <= 0 || cx0 + CC - 1 >= a.wc - 1;
            const int j = tid & (CC - 1);
            const int j_own = min(4 + SW / 2, a.wc - cx0);  // owned ring columns: [4, j_own)
            const bool own_j = j >= 4 && j < j_own;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int vi = (tid >> 5) + 4 * pass, v = vi >> 2, i = vi & 3;
                const int cy = cyn + i;
                const float4 *ev = &sm.vp[v][i][j], *od = &sm.vp[v][i][CVVDP_BF_VP + j];  // staged columns 2j+{0,2,4} / {1,3}
                float4 s = K0 * ev[0];
                s = fma4(K1, od[0], s);
                s = fma4(K2, ev[1], s);
                s = fma4(K1, od[1], s);
                s = fma4(K0, ev[2], s);
                if (edge_x) {
                    const int cx = cx0 + j;
                    auto vp_at = [&](int f) -> const float4 & { return sm.vp[v][i][(f & 1) * CVVDP_BF_VP + (f >> 1)]; };
                    if (cx == 0) {  // l.205
                        s = fma4(K1, vp_at(0 - fx0), s);
                        s = fma4(K0, vp_at(min(1, a.w - 1) - fx0), s);
                    }
                    if (cx == a.wc - 1) {  // l.206-209: the parity of the ROW count chooses the rule
                        if (rows_odd) {
                            s = fma4(K1, vp_at(a.w - 1 - fx0), s);
                            s = fma4(K0, vp_at(max(a.w - 2, 0) - fx0), s);
                        } else {
                            s = fma4(K0, vp_at(a.w - 1 - fx0), s);
                        }
                    }
                }
                sm.crs[v][(cy - cbase) & 7][j] = s;
                if (own_j && cy >= own_cy0 && cy < own_cy1)
                    crs_out[(unsigned)(v * (int)ncpix + cy * a.wc + cx0 + j)] = s;
            }
        }
        __syncthreads();  // coarse ring complete; vp (= mm) is free
        // ---- phase A: one 2x2 quad per thread: expand, contrast, CSF -> mm, df ----
        if (do_a) {
            const int gy = a0 + 2 * qy;
            if (gy < a_end && a_cols) {
                float4 e[2][4];
                // coarse rows a0/2-1+qy+{0,1,2}, columns cx0+qx+{0,1,2}, clamped to the coarse image (replicate padding)
                int rs[3], cs[3];
                if (a0 / 2 - 1 < 0 || a0 / 2 + 4 > a.hc - 1 || cx0 < 0 || cx0 + CC - 1 > a.wc - 1) {  // uniform: at a border
#pragma unroll
                    for (int r = 0; r < 3; ++r) rs[r] = (min(max(a0 / 2 - 1 + qy + r, 0), a.hc - 1) - cbase) & 7;
#pragma unroll
                    for (int c = 0; c < 3; ++c) cs[c] = min(max(cx0 + qx + c, 0), a.wc - 1) - cx0;
                } else {
#pragma unroll
                    for (int r = 0; r < 3; ++r) rs[r] = (a0 / 2 - 1 + qy + r - cbase) & 7;
#pragma unroll
                    for (int c = 0; c < 3; ++c) cs[c] = qx + c;
                }
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    float4 ve[3], vo[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float4 c0 = sm.crs[v][rs[0]][cs[c]], c1 = sm.crs[v][rs[1]][cs[c]], c2 = sm.crs[v][rs[2]][cs[c]];
                        ve[c] = fma4(0.1f, c2, fma4(0.8f, c1, 0.1f * c0));
                        vo[c] = fma4(0.5f, c2, 0.5f * c1);
                    }
                    e[v][0] = fma4(wA2, ve[2], fma4(wA1, ve[1], wA0 * ve[0]));
                    e[v][1] = fma4(wB2, ve[2], fma4(wB1, ve[1], wB0 * ve[0]));
                    e[v][2] = fma4(wA2, vo[2], fma4(wA1, vo[1], wA0 * vo[0]));
                    e[v][3] = fma4(wB2, vo[2], fma4(wB1, vo[1], wB0 * vo[0]));
                }
                float4 mm[4], df[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int ry = 2 * qy + (k >> 1), rx = 2 * qx + ((k & 1) ^ sw_odd);
                    band_pixel<false>(a, sm.lut, sm.fine_a[0][ry][rx + 4], sm.fine_a[1][ry][rx + 4], e[0][k], e[1][k], mm[k], df[k]);
                }
                const int ix = ex0 + 2 * qx - x0;
                const bool in_strip = ix >= 0 && ix < SW;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int ry = 2 * qy + (k >> 1), cofs = (k & 1) ^ sw_odd, rx = 2 * qx + cofs;
                    sm.mm[ry][rx] = mm[k];
                    if (in_strip) sm.df[(a0 + ry) & (CVVDP_B2_DFR - 1)][ix + cofs] = df[k];
                }
            }
        }
        __syncthreads();
        // ---- prefetch the next step's stage while phases B and C run ----
        if (a0 + CVVDP_B2_RB < a_end) stage(a0 + CVVDP_B2_RB);
        // ---- phase B: horizontal pass of the phase-uncertainty Gaussian for the new rows ----
        if (do_a && tid < B2Geom<EW>::B_TASKS) {
            const int gy = a0 + b_r, gxb = x0 + b_xg * 4;
            if (gy < a_end && gxb < a.w) {
                float4 win[2 * CVVDP_BHALO + 4];
                if (x_edge) {
#pragma unroll
                    for (int j = 0; j < 2 * CVVDP_BHALO + 4; ++j) {
                        int lx = reflect_idx(gxb + j - CVVDP_BHALO, a.w) - ex0;
                        lx = min(max(lx, 0), EW - 1);
                        win[j] = sm.mm[b_r][lx];
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 2 * CVVDP_BHALO + 4; ++j) win[j] = sm.mm[b_r][b_xg * 4 + j];
                }
                float4 *dst = &sm.hb[hslot0 + b_r][b_xg * 4];
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    float4 s = f4(0.f);
#pragma unroll
                    for (int k = 0; k < 2 * CVVDP_BHALO + 1; ++k) s = fma4(a.kern[k], win[o + k], s);
                    dst[o] = s;
                }
            }
        }
        __syncthreads();
        // ---- phase C: vertical pass, masking, clamp, pooling for rows [a0-6, a0+2) ----
        if (a0 >= y_begin && tid < B2Geom<EW>::C_TASKS) {
            const int gx = x0 + c_ix;
            const int cyb = a0 - hal + c_rg * 4;
            if (gx < a.w && cyb + 3 >= ys && cyb < ye) {
                float4 win[2 * CVVDP_BHALO + 4];
                const bool y_edge = (cyb - CVVDP_BHALO < 0) || (cyb + 3 + CVVDP_BHALO >= a.h);
                if (y_edge) {
#pragma unroll
                    for (int j = 0; j < 2 * CVVDP_BHALO + 4; ++j) {
                        int yy = reflect_idx(cyb + j - CVVDP_BHALO, a.h);
                        yy = min(max(yy, y_begin), a.h - 1);
                        win[j] = sm.hb[(yy - y_begin) % CVVDP_BF_HBR][c_ix];
                    }
                } else {  // row cyb-6 = a0-12+4 c_rg sits 12+4 c_rg slots after the slot of row a0 (mod 24)
                    int s0 = hslot0 + 12 + 4 * c_rg;
                    s0 -= s0 >= CVVDP_BF_HBR ? CVVDP_BF_HBR : 0;
#pragma unroll
                    for (int j = 0; j < 2 * CVVDP_BHALO + 4; ++j) {
                        int sl = s0 + j;
                        sl -= sl >= CVVDP_BF_HBR ? CVVDP_BF_HBR : 0;
                        win[j] = sm.hb[sl][c_ix];
                    }
                }
                float4 D[4];
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const int gy = cyb + o;
                    float4 m = f4(0.f);
#pragma unroll
                    for (int k = 0; k < 2 * CVVDP_BHALO + 1; ++k) m = fma4(a.kern[k], win[o + k], m);
                    D[o] = band_mask(a, m, sm.df[gy & (CVVDP_B2_DFR - 1)][c_ix], eps_q, eps_p);
                }
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const int gy = cyb + o;
                    const bool live = gy >= ys && gy < ye;
                    float4 t;
                    if (BETA2) {
                        const float e2 = 2.f * a.eps;
                        t = make_float4(D[o].x * (D[o].x + e2), D[o].y * (D[o].y + e2), D[o].z * (D[o].z + e2), D[o].w * (D[o].w + e2));
                    } else {
                        t = make_float4(f_pow(D[o].x + a.eps, a.beta) - eps_b, f_pow(D[o].y + a.eps, a.beta) - eps_b,
                                        f_pow(D[o].z + a.eps, a.beta) - eps_b, f_pow(D[o].w + a.eps, a.beta) - eps_b);
                    }
                    acc.x += live ? t.x : 0.f;
                    acc.y += live ? t.y : 0.f;
                    acc.z += live ? t.z : 0.f;
                    acc.w += live ? t.w : 0.f;
                    if (HM && live) {
                        const float eb = f_pow(a.eps, a.hm_beta);
                        float s = (f_pow(D[o].x * a.hm_w[0] + a.eps, a.hm_beta) - eb) + (f_pow(D[o].y * a.hm_w[1] + a.eps, a.hm_beta) - eb) +
                                  (f_pow(D[o].z * a.hm_w[2] + a.eps, a.hm_beta) - eb) + (f_pow(D[o].w * a.hm_w[3] + a.eps, a.hm_beta) - eb);
                        const float ib = 1.f / a.hm_beta;
                        a.hm[(long long)pair * npix + (long long)gy * a.w + gx] = (f_pow(s + a.eps, ib) - f_pow(a.eps, ib)) * a.hm_scale;
                    }
                }
            }
        }
        if (a0 >= y_begin) {
            hslot0 += CVVDP_B2_RB;
            hslot0 -= hslot0 >= CVVDP_BF_HBR ? CVVDP_BF_HBR : 0;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
        acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if ((tid & 31) == 0) {
        sm.red[tid >> 5][0] = acc.x;
        sm.red[tid >> 5][1] = acc.y;
        sm.red[tid >> 5][2] = acc.z;
        sm.red[tid >> 5][3] = acc.w;
    }
    __syncthreads();
    if (tid < 4) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < CVVDP_B2_THREADS / 32; ++w) s += sm.red[w][tid];
        const int tile = blockIdx.y * gridDim.x + blockIdx.x, ntiles = gridDim.x * gridDim.y;
        a.partials[((long long)pair * ntiles + tile) * 4 + tid] = s;
    }
}

// =================================================================================================
// Baseband (cvvdp_metric.py:711-712 with lpyr_dec.py:378-384): L_bkg is the spatial mean of the
// clamped achromatic plane of each image; D = |T - R| * S (no gain, masking or clamp).
// One CTA per (batch item, frame).
// =================================================================================================
struct BasebandArgs {
    const float4 *g;       // level L-1 [pairs*2][npix]
    const float4 *lut;     // [32] rows, pre-scaled: row*log2(10)+log2(sens)
    float *partials;       // [pairs][1][4]
    float *hm;             // [pairs][npix] or null
    float4 *feat;          // feature mode: [3][pairs][npix] planes |T|S, |R|S, D, or null
    int npix;
    float lut_a, lut_b;
    float eps, beta;
    float hm_w[4];
    float hm_beta;
};

__device__ __forceinline__ float block_sum_256(float v, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    return s;
}

__global__ void __launch_bounds__(256) k_baseband(const BasebandArgs a) {
    __shared__ float red[8];
    const int tid = threadIdx.x, pair = blockIdx.x;
    const float4 *gt = a.g + (long long)pair * 2 * a.npix, *gr = gt + a.npix;
    float st = 0.f, sr = 0.f;
    for (int i = tid; i < a.npix; i += 256) {
        st += fmaxf(gt[i].x, 0.01f);
        sr += fmaxf(gr[i].x, 0.01f);
    }
    const float Lt = block_sum_256(st, red) / (float)a.npix;
    const float Lr = block_sum_256(sr, red) / (float)a.npix;
    float ind = fminf(fmaxf(fmaf(log2f(Lr), a.lut_a, a.lut_b), 0.f), (float)(CVVDP_CSF_LUT_N - 1));
    const int i0 = (int)ind;
    const float fr = ind - (float)i0;
    const int i1 = min(i0 + 1, CVVDP_CSF_LUT_N - 1);
    const float4 va = a.lut[i0], vb = a.lut[i1];
    const float S[4] = {exp2f(va.x * (1.f - fr) + vb.x * fr), exp2f(va.y * (1.f - fr) + vb.y * fr),
                        exp2f(va.z * (1.f - fr) + vb.z * fr), exp2f(va.w * (1.f - fr) + vb.w * fr)};
    const float eps_b = powf(a.eps, a.beta);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = tid; i < a.npix; i += 256) {
        const float4 t = gt[i], r = gr[i];
        const float tv[4] = {t.x, t.y, t.z, t.w}, rv[4] = {r.x, r.y, r.z, r.w};
        float D[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float ct = fminf(tv[c] / Lt, 1000.f), cr = fminf(rv[c] / Lr, 1000.f);
            D[c] = fabsf(ct - cr) * S[c];
            acc[c] += (a.beta == 2.0f) ? D[c] * (D[c] + 2.f * a.eps) : powf(D[c] + a.eps, a.beta) - eps_b;
        }
        if (a.hm) {
            const float eb = powf(a.eps, a.hm_beta);
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) s += powf(D[c] * a.hm_w[c] + a.eps, a.hm_beta) - eb;
            const float ib = 1.f / a.hm_beta;
            a.hm[(long long)pair * a.npix + i] = powf(s + a.eps, ib) - powf(a.eps, ib);
        }
        if (a.feat) {  // cvvdp_ml_metric.py:352 on the baseband: |T_f| S, |R_f| S, D
            const long long plane = (long long)gridDim.x * a.npix, off = (long long)pair * a.npix + i;
            float ft[4], fr[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                ft[c] = fabsf(fminf(tv[c] / Lt, 1000.f)) * S[c];
                fr[c] = fabsf(fminf(rv[c] / Lr, 1000.f)) * S[c];
            }
            a.feat[off] = make_float4(ft[0], ft[1], ft[2], ft[3]);
            a.feat[plane + off] = make_float4(fr[0], fr[1], fr[2], fr[3]);
            a.feat[2 * plane + off] = make_float4(D[0], D[1], D[2], D[3]);
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float s = block_sum_256(acc[c], red);
        if (tid == 0) a.partials[(long long)pair * 4 + c] = s;
    }
}

// =================================================================================================
// Spatial pooling epilogue: fixed-order sum of the per-tile partials, then
// Q = safe_pow(sum / N, 1 / beta)  (lp_norm, cvvdp_metric.py:1032-1048) -> Q_per_ch[B, C, F, L].
// One warp per (pair, band).
// =================================================================================================
struct FinalizeArgs {
    const float *partials[CVVDP_MAX_BANDS];
    int ntiles[CVVDP_MAX_BANDS];
    int npix[CVVDP_MAX_BANDS];
    int L, C, B, n, f_off, F_total;
    float beta, eps;
    float *Q;  // [B][C][F_total][L]
};

__global__ void __launch_bounds__(128) k_finalize(const FinalizeArgs a) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int total = a.B * a.n * a.L;
    const bool live = warp < total;  // whole warps only: keep every lane in the shuffles
    const int wi = live ? warp : 0;
    const int band = wi % a.L, pair = wi / a.L;
    const int nt = a.ntiles[band];
    const float4 *src = reinterpret_cast<const float4 *>(a.partials[band]) + (long long)pair * nt;
    float4 acc = f4(0.f);
    for (int t = lane; t < nt; t += 32) acc = acc + src[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
        acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (live && lane < a.C) {
        const float s = lane == 0 ? acc.x : (lane == 1 ? acc.y : (lane == 2 ? acc.z : acc.w));
        const float ib = 1.f / a.beta;
        const float q = powf(s / (float)a.npix[band] + a.eps, ib) - powf(a.eps, ib);
        const int b = pair / a.n, f = pair - b * a.n;
        a.Q[(((long long)b * a.C + lane) * a.F_total + (a.f_off + f)) * a.L + band] = q;
    }
}

// =================================================================================================
// do_pooling_and_jods + met2jod (cvvdp_metric.py:610-658).  One CTA per batch item.
// =================================================================================================
struct PoolArgs {
    const float *Q;  // [B][C][F][L]
    float *jod;      // [B]
    int B, C, F, L;
    float ch_w[4], bb_w[4];
    float beta_sch, beta_tch, beta_t, image_int, jod_a, jod_exp, eps;
};
__device__ __forceinline__ float spow_acc(float x, float p, float eps) { return powf(x + eps, p) - powf(eps, p); }
__device__ __forceinline__ float met2jod_dev(float Q, float jod_a, float jod_exp) {
    const float Qt = 0.1f;
    if (Q <= Qt) return 10.f - jod_a * powf(Qt, jod_exp - 1.f) * Q;
    return 10.f - jod_a * powf(Q, jod_exp);
}
__global__ void __launch_bounds__(256) k_pool(const PoolArgs a) {
    __shared__ float red[8];
    const int b = blockIdx.x, tid = threadIdx.x;
    float acc = 0.f, q_img = 0.f;
    for (int f = tid; f < a.F; f += 256) {
        float s_tc = 0.f;
        for (int c = 0; c < a.C; ++c) {
            float s = 0.f;
            for (int l = 0; l < a.L; ++l) {
                // reference order: Q * per_ch_w * per_sband_w (l.625)
                float v = a.Q[(((long long)b * a.C + c) * a.F + f) * a.L + l] * a.ch_w[c];
                v = (l == a.L - 1) ? v * a.bb_w[c] : v;
                s += spow_acc(v, a.beta_sch, a.eps);
            }
            const float q_sc = spow_acc(s, 1.f / a.beta_sch, a.eps);
            s_tc += spow_acc(q_sc, a.beta_tch, a.eps);
        }
        const float q_tc = spow_acc(s_tc, 1.f / a.beta_tch, a.eps);
        q_img = q_tc;
        acc += spow_acc(q_tc, a.beta_t, a.eps);
    }
    const float tot = block_sum_256(acc, red);
    if (tid == 0) {
        float Q;
        // l.636: images (one frame, three channels -- a one-frame slice of a video keeps its transient channel
        // and the temporal pooling formula)
        if (a.F == 1 && a.C == 3) Q = q_img * a.image_int;
        else Q = spow_acc(tot / (float)a.F, 1.f / a.beta_t, a.eps);  // l.638
        a.jod[b] = met2jod_dev(Q, a.jod_a, a.jod_exp);
    }
}

// =================================================================================================
// Feature pooling for the ML heads (cvvdp_feature_pooling, cvvdp_ml_metric.py:78-106): mean and
// variance of |T|S, |R|S and D over feature_size x feature_size patches of one band
// (AvgPool2d(ceil_mode=True): a ragged border patch is averaged over the pixels it covers).
// One CTA per (patch, item-frame); fixed-order reduction.  Output [B][F][ph][pw][C][6].
// =================================================================================================
struct FeaturePoolArgs {
    const float4 *feat;   // [3][pairs][h*w]
    long long feat_plane;
    float *out;           // this band: [B][F_total][ph][pw][C][6]
    int h, w, ps, ph, pw;
    int C, n, f_off, F_total;
};
__global__ void __launch_bounds__(128) k_feature_pool(const FeaturePoolArgs a) {
    __shared__ float red[4][24];
    const int tid = threadIdx.x, pxi = blockIdx.x, pyi = blockIdx.y, pair = blockIdx.z;
    const int x0 = pxi * a.ps, y0 = pyi * a.ps;
    const int pwid = min(a.ps, a.w - x0), phgt = min(a.ps, a.h - y0);
    const long long npix = (long long)a.h * a.w;
    float s[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = 0.f;
    for (int i = tid; i < pwid * phgt; i += 128) {
        const int r = i / pwid, c = i - r * pwid;
        const long long off = (long long)pair * npix + (long long)(y0 + r) * a.w + (x0 + c);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float4 v = a.feat[k * a.feat_plane + off];
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                s[k * 8 + ch] += vv[ch];
                s[k * 8 + 4 + ch] = fmaf(vv[ch], vv[ch], s[k * 8 + 4 + ch]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 24; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
    }
    if ((tid & 31) == 0) {
#pragma unroll
        for (int i = 0; i < 24; ++i) red[tid >> 5][i] = s[i];
    }
    __syncthreads();
    if (tid < 3 * a.C) {
        const int k = tid / a.C, ch = tid - k * a.C;
        const float cnt = (float)(pwid * phgt);
        const float sum = (red[0][k * 8 + ch] + red[1][k * 8 + ch]) + (red[2][k * 8 + ch] + red[3][k * 8 + ch]);
        const float sq = (red[0][k * 8 + 4 + ch] + red[1][k * 8 + 4 + ch]) + (red[2][k * 8 + 4 + ch] + red[3][k * 8 + 4 + ch]);
        const float mean = sum / cnt;
        const int b = pair / a.n, f = pair - b * a.n;
        float *dst = a.out + (((((long long)b * a.F_total + (a.f_off + f)) * a.ph + pyi) * a.pw + pxi) * a.C + ch) * 6 + 2 * k;
        dst[0] = mean;
        dst[1] = sq / cnt - mean * mean;
    }
}

// =================================================================================================
// Heat map: reconstruct the per-band difference planes (lpyr_dec.py:328-335), then
// 1 - met2jod(.)/10 in fp16 (cvvdp_metric.py:743-744, 398).
// =================================================================================================
struct ExpandAddArgs {
    const float *coarse;  // [planes][hc*wc]
    float *fine;          // [planes][h*w]   fine += expand(coarse)
    int h, w, hc, wc;
};
__device__ __forceinline__ float expand_at(const float *c, int hc, int wc, int y, int x) {
    const int j = y >> 1, i = x >> 1;
    float col[3];
    const int jm = max(j - 1, 0), jp = min(j + 1, hc - 1), im = max(i - 1, 0), ip = min(i + 1, wc - 1);
    const int xs[3] = {im, i, ip};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float c0 = c[jm * wc + xs[k]], c1 = c[j * wc + xs[k]], c2 = c[jp * wc + xs[k]];
        col[k] = (y & 1) ? fmaf(0.5f, c2, 0.5f * c1) : fmaf(0.1f, c2, fmaf(0.8f, c1, 0.1f * c0));
    }
    return (x & 1) ? fmaf(0.5f, col[2], 0.5f * col[1]) : fmaf(0.1f, col[2], fmaf(0.8f, col[1], 0.1f * col[0]));
}
__global__ void __launch_bounds__(256) k_expand_add(const ExpandAddArgs a) {
    const long long npix = (long long)a.h * a.w;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int plane = blockIdx.y;
    if (p >= npix) return;
    const int y = (int)(p / a.w), x = (int)(p - (long long)y * a.w);
    const float *c = a.coarse + (long long)plane * a.hc * a.wc;
    a.fine[plane * npix + p] += expand_at(c, a.hc, a.wc, y, x);
}
struct HeatmapOutArgs {
    const float *img;     // [B*n][npix]   (B == 1)
    unsigned short *out;  // fp16 [F_total][npix]
    long long npix;
    int f_off;
    float jod_a, jod_exp;
};
__global__ void __launch_bounds__(256) k_heatmap_out(const HeatmapOutArgs a) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (p >= a.npix) return;
    const float v = 1.f - met2jod_dev(a.img[f * a.npix + p], a.jod_a, a.jod_exp) / 10.f;
    a.out[(long long)(a.f_off + f) * a.npix + p] = float_to_half_bits(v);
}


// =================================================================================================
// Coloured heat maps ('threshold' / 'supra-threshold'): visualize_diff_map + vis_tonemap
// (visualize_diff_map.py:23-106) on top of the reconstructed difference map.  The context image is the TEST
// sustained achromatic channel of the block, R[:,0] (cvvdp_metric.py:399-401) = the .x lane of the level-0 test
// planes, tone-mapped with the statistics of the BLOCK of frames the reference visualises at once: minimum
// positive value, maximum, a 1024-bin histogram of the log image -> cube-root histogram equalisation.
// Four small kernels: min/max, histogram, curve (one CTA), colour.
// =================================================================================================
#define CVVDP_HM_BINS 1024
struct HmToneArgs {
    const float4 *lv0;     // level 0: [n][2][npix]; plane 2 f = test video of frame f (heat maps need a batch of one)
    long long npix;
    int n;
    unsigned *minmax;      // [0] = bits of the smallest positive context value, [1] = bits of the largest
    int *hist;             // [CVVDP_HM_BINS]
    float *curve;          // [0..1023] tone curve v, [1024..2047] b_scale, [2048] b_min, [2049] b_max
    float dr;              // 0.6
};
__global__ void __launch_bounds__(256) k_hm_minmax(const HmToneArgs a) {
    const float4 *src = a.lv0 + (long long)blockIdx.y * 2 * a.npix;
    unsigned mn = 0x7f800000u, mx = 0u;  // positive floats order like their bit patterns
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < a.npix; p += (long long)gridDim.x * 256) {
        const float y = src[p].x;
        if (y > 0.f) {
#ifdef __CUDA_ARCH__
            const unsigned b = __float_as_uint(y);
#else
            unsigned b;
            memcpy(&b, &y, 4);
#endif
            mn = min(mn, b);
            mx = max(mx, b);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&a.minmax[0], mn);
        atomicMax(&a.minmax[1], mx);
    }
}
__device__ __forceinline__ float hm_log_context(float y, float clampval) { return logf(fmaxf(y, clampval)); }
__global__ void __launch_bounds__(256) k_hm_hist(const HmToneArgs a) {
    __shared__ int sh[CVVDP_HM_BINS];
    for (int i = threadIdx.x; i < CVVDP_HM_BINS; i += 256) sh[i] = 0;
    __syncthreads();
    const float clampval = bits_as_float(a.minmax[0]);
    const float b_min = logf(clampval), b_max = logf(bits_as_float(a.minmax[1]));
    const float4 *src = a.lv0 + (long long)blockIdx.y * 2 * a.npix;
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < a.npix; p += (long long)gridDim.x * 256) {
        const float b = hm_log_context(src[p].x, clampval);
        // torch.histc (CUDA): bin = (int)((x - min) * bins / (max - min)), the maximum goes to the last bin
        int bin = (int)((b - b_min) * (float)CVVDP_HM_BINS / (b_max - b_min));
        bin = min(max(bin, 0), CVVDP_HM_BINS - 1);
        atomicAdd(&sh[bin], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CVVDP_HM_BINS; i += 256)
        if (sh[i]) atomicAdd(&a.hist[i], sh[i]);
}
// One CTA of 1024 threads: b_p = hist / sum; dy = b_p^(1/3) / sum(b_p^(1/3)); v = cumsum(dy) dr + (1 - dr) / 2;
// b_scale = linspace(b_min, b_max, 1024)  (visualize_diff_map.py:34-43).
__global__ void __launch_bounds__(CVVDP_HM_BINS) k_hm_curve(const HmToneArgs a) {
    __shared__ float s_scan[CVVDP_HM_BINS];
    __shared__ float s_red[32];
    const int i = threadIdx.x;
    const float clampval = bits_as_float(a.minmax[0]);
    const float b_min = logf(clampval), b_max = logf(bits_as_float(a.minmax[1]));
    const float total = (float)a.npix * (float)a.n;
    const float bp = (float)a.hist[i] / total;
    const float r = powf(bp, (float)(1.0 / 3.0));
    // sum of r over the CTA
    float t = r;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((i & 31) == 0) s_red[i >> 5] = t;
    __syncthreads();
    float sum = 0.f;
    for (int w = 0; w < 32; ++w) sum += s_red[w];
    // inclusive scan of dy (Hillis-Steele over shared memory; 1024 elements)
    s_scan[i] = r / sum;
    __syncthreads();
    for (int o = 1; o < CVVDP_HM_BINS; o <<= 1) {
        const float add = i >= o ? s_scan[i - o] : 0.f;
        __syncthreads();
        s_scan[i] += add;
        __syncthreads();
    }
    a.curve[i] = s_scan[i] * a.dr + (1.0f - a.dr) / 2.0f;
    // torch.linspace: start + i step for the first half, end - (steps - 1 - i) step for the second
    const float step = (b_max - b_min) / (float)(CVVDP_HM_BINS - 1);
    a.curve[CVVDP_HM_BINS + i] = i < CVVDP_HM_BINS / 2 ? b_min + step * (float)i : b_max - step * (float)(CVVDP_HM_BINS - 1 - i);
    if (i == 0) {
        a.curve[2 * CVVDP_HM_BINS] = b_min;
        a.curve[2 * CVVDP_HM_BINS + 1] = b_max;
    }
}
struct HmColourArgs {
    const float *img;      // reconstructed difference map [n][npix]
    const float4 *lv0;     // context, as in HmToneArgs
    const unsigned *minmax;
    const float *curve;
    unsigned short *out;   // fp16 [3][F_total][npix]
    long long npix;
    int f_off, F_total;
    float jod_a, jod_exp, dr;
    int n_map;             // colour map entries (visualize_diff_map.py:57-84)
    float map_in[5];
    float map_ch[5][3];    // colour / (luminance + 1e-4)
};
// interp1 (interp.py:22-31, 81-89): imax = bucketize(x, xs) (first i with xs[i] >= x), clamped; imin = imax - 1
__device__ __forceinline__ float hm_interp1(const float *xs, const float *vs, int stride, int n, float x) {
    int lo = 0, hi = n;  // first index with xs[i] >= x
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (xs[mid] < x) lo = mid + 1;
        else hi = mid;
    }
    const int imax = min(lo, n - 1), imin = max(imax - 1, 0);
    float fr = (x - xs[imin]) / (xs[imax] - xs[imin] + 0.000001f);
    if (imax == imin || fr < 0.f) fr = 0.f;
    return vs[imin * stride] * (1.0f - fr) + vs[imax * stride] * fr;
}
__global__ void __launch_bounds__(256) k_hm_colour(const HmColourArgs a) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (p >= a.npix) return;
    float d = 1.f - met2jod_dev(a.img[f * a.npix + p], a.jod_a, a.jod_exp) / 10.f;
    d = fminf(fmaxf(d, 0.f), 1.f);
    const float clampval = bits_as_float(a.minmax[0]);
    const float b = hm_log_context(a.lv0[(long long)f * 2 * a.npix + p].x, clampval);
    const float b_min = a.curve[2 * CVVDP_HM_BINS], b_max = a.curve[2 * CVVDP_HM_BINS + 1];
    float tmo;
    if (b_max - b_min < a.dr) tmo = (b - b_min) / (b_max - b_min + 1e-3f) * a.dr + (1.f - a.dr) / 2.f;  // l.30-32
    else tmo = hm_interp1(a.curve + CVVDP_HM_BINS, a.curve, 1, CVVDP_HM_BINS, b);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // the reference stores the colour in fp16 before it multiplies by the fp32 tone-mapped context (l.98-104)
        const float col = half_bits_to_float(float_to_half_bits(hm_interp1(a.map_in, &a.map_ch[0][c], 3, a.n_map, d)));
        const float v = fminf(fmaxf(col * tmo, 0.f), 1.f);
        a.out[((long long)c * a.F_total + a.f_off + f) * a.npix + p] = float_to_half_bits(v);
    }
}

}  // namespace cvvdp
