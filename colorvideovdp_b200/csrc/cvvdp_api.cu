// cvvdp_api.cu -- host side of libcvvdp_b200.so: context, per-clip planning, kernel launches and the
// host-buffer streaming path behind the C ABI declared in include/cvvdp_b200.h.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "cvvdp_kernels.cuh"

static const size_t kFlagsBytes = (4 + CVVDP_MEAN_SLOTS) * sizeof(int);  // validation counters + first-frame partial sums

using namespace cvvdp;

// Band kernel configuration (B2Geom / k_band2): 48-column strips, conflict-free phase A, phase C trailing by 16 rows.
#define CVVDP_BAND_EW 60
#define CVVDP_BAND_CF true
#define CVVDP_BAND_LAG true

namespace {

thread_local std::string g_create_error;

struct LevelBuf {
    int h = 0, w = 0;
    float4 *g = nullptr;        // [B*nb*2][h*w]
    float *partials = nullptr;  // [B*nb][tiles][4]
    float *hm = nullptr;        // [nb][h*w] (heat map only)
    float4 *feat = nullptr;     // [3][B*nb][h*w] (feature mode only): |T|S, |R|S, D
    int ph = 0, pw = 0;         // feature patches of this band
    long long feat_off = 0;     // float offset of this band's tensor in the caller's feature buffer
    float4 *lut = nullptr;      // [32]
    int tiles_x = 0, tiles_y = 0;  // band kernel grid: column strips x row segments
    int seg_rows = 0;
    int do_blur = 0;
    TensorMap3D tm;  // fp32 view [planes][h][4w] of g; box depends on the role (see make_tensor_map)
    TensorMap3D tm_as_coarse;
    TensorMap3D tm_reduce_in;  // box {63 px, 19 rows, 1 plane}
    TensorMap3D tm_fa, tm_fb;  // fused band kernel: boxes {64 px, 11 rows, 2 planes} and {4 px, 11 rows, 2 planes}
    bool fused = false;        // level i+1 is computed by this level's band kernel (k_band2f), no reduce launch
    bool tm_ok = false;
};

struct Staging {
    void *buf[2] = {nullptr, nullptr};  // test, ref
    size_t bytes = 0;
    cudaEvent_t copied = nullptr, consumed = nullptr;
};

}  // namespace

struct cvvdp_b200_ctx {
    int device = 0;
    cvvdp_b200_params P;
    cvvdp_b200_csf_lut lut;
    cvvdp_b200_display disp;
    bool have_display = false, planned = false;
    cvvdp_b200_job job;
    cvvdp_b200_plan_info info;
    std::vector<LevelBuf> lv;
    void *arena = nullptr;
    size_t arena_bytes = 0;
    float blur_kern[2 * CVVDP_BHALO + 1];
    int blur_pad = 0;
    int max_smem_optin = 0;
    int num_sms = 148;
    cudaStream_t copy_stream = nullptr, work_stream = nullptr;
    Staging stage[2];
    float *feat_out = nullptr;    // caller's device buffer for the feature tensors (feature mode)
    int feature_size = 0;         // ceil(ppd)
    long long feat_total = 0;     // floats in the feature buffer
    float *q_dev = nullptr;       // for process_host / pool
    size_t q_dev_bytes = 0;
    void *hm_dev = nullptr;
    size_t hm_dev_bytes = 0;
    // pinned bounce buffers for PAGEABLE host clips (see upload())
    static const int kPinSlots = 3;
    void *pin_buf[kPinSlots] = {nullptr, nullptr, nullptr};
    cudaEvent_t pin_done[kPinSlots] = {nullptr, nullptr, nullptr};
    size_t pin_bytes = 0;
    int pin_next = 0;
    unsigned *hm_tone_dev = nullptr;  // coloured heat maps: [0..1] min/max bits, [2..1025] histogram, then 2050 floats of tone curve
    cudaStream_t d2h_stream = nullptr;
    cudaEvent_t dev_done = nullptr;   // end of the last process_device / pool_device on the caller's stream
    bool dev_pending = false;
    cudaEvent_t hm_ready = nullptr, hm_copied = nullptr;
    int *flags_dev = nullptr;     // [0..2] input validation counters; then CVVDP_MEAN_SLOTS floats: partial DKL-A sums of test frame 0
    long long launches = 0;
    bool prof = false;
    struct ProfRec {
        int kind, level;
        double bytes;
        cudaEvent_t a, b;
    };
    std::vector<ProfRec> prof_recs;
    std::string err;
};

namespace {

int fail(cvvdp_b200_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    else g_create_error = buf;
    return code;
}

#define CU_CHECK(ctx, call)                                                                         \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(ctx, CVVDP_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                        \
    } while (0)

// Every ABI entry point runs on the context's device and leaves the caller's current device as it found it
// (PyTorch reads the runtime's current device: a library that changes it behind torch's back redirects the caller's
// next allocation).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) err = cudaSetDevice(device);
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Tiled tensor map over `planes` float4 planes of h x w pixels, viewed as fp32 [planes][h][4w].
bool make_tensor_map(TensorMap3D *m, const float4 *base, int w, int h, int planes, int box_px, int box_rows, int box_planes = 2,
                     int l2_promotion = 128) {
#ifdef CVVDP_EMU
    m->base = (const float *)base;
    m->dim[0] = 4 * w;
    m->dim[1] = h;
    m->dim[2] = planes;
    m->stride[0] = 1;
    m->stride[1] = 4LL * w;
    m->stride[2] = 4LL * w * h;
    m->box[0] = 4 * box_px;
    m->box[1] = box_rows;
    m->box[2] = box_planes;
    (void)l2_promotion;
    return true;
#else
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return false;
        encode = (EncodeFn)fn;
    }
    if (4 * box_px > 256 || box_rows > 256) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)4 * w, (cuuint64_t)h, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)w * 16, (cuuint64_t)w * h * 16};  // bytes, dims 1 and 2
    const cuuint32_t box[3] = {(cuuint32_t)(4 * box_px), (cuuint32_t)box_rows, (cuuint32_t)box_planes};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapL2promotion promo = l2_promotion == 0    ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                         : l2_promotion == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                         : l2_promotion == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                                               : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
#endif
}

// Brackets one kernel launch with events when profiling is on; always counts the launch.
struct LaunchScope {
    cvvdp_b200_ctx *ctx;
    cudaStream_t st;
    cudaEvent_t b = nullptr;
    LaunchScope(cvvdp_b200_ctx *c, cudaStream_t s, int kind, int level, double bytes) : ctx(c), st(s) {
        ctx->launches++;
        if (!ctx->prof) return;
        cvvdp_b200_ctx::ProfRec r;
        r.kind = kind;
        r.level = level;
        r.bytes = bytes;
        cudaEventCreateWithFlags(&r.a, 0);
        cudaEventCreateWithFlags(&r.b, 0);
        cudaEventRecord(r.a, st);
        b = r.b;
        ctx->prof_recs.push_back(r);
    }
    ~LaunchScope() {
        if (b) cudaEventRecord(b, st);
    }
};

void free_plan(cvvdp_b200_ctx *ctx) {
    if (ctx->arena) cudaFree(ctx->arena);
    ctx->arena = nullptr;
    ctx->arena_bytes = 0;
    ctx->lv.clear();
    for (auto &s : ctx->stage) {
        for (auto &b : s.buf) {
            if (b) cudaFree(b);
            b = nullptr;
        }
        s.bytes = 0;
    }
    ctx->planned = false;
}

// ---- host restatements of the per-clip constants -------------------------------------------------

// lpyr_dec.py:18-52 (+ cvvdp_metric.py:685-686)
int band_setup(int W, int H, double ppd, cvvdp_b200_plan_info *info) {
    const int max_levels = (int)floor(log2((double)std::min(H, W))) - 1;
    int max_band = max_levels;
    for (int k = 0; k < 15; ++k) {
        const double band = (k == 0 ? 1.0 : 0.3228 * pow(2.0, -(double)(k - 1))) * ppd / 2.0;
        if (band <= 0.2) {
            max_band = k;
            break;
        }
    }
    int height = std::max(0, std::min(max_band + 1, max_levels));
    if (height + 1 > CVVDP_MAX_BANDS) height = CVVDP_MAX_BANDS - 1;
    info->n_bands = height + 1;
    int h = H, w = W;
    for (int i = 0; i <= height; ++i) {
        info->rho_band[i] = (float)((i == 0 ? 1.0 : 0.3228 * pow(2.0, -(double)(i - 1))) * ppd / 2.0);
        info->band_height[i] = h;
        info->band_width[i] = w;
        h = (h + 1) / 2;
        w = (w + 1) / 2;
    }
    info->rho_band[height] = 0.1f;
    return height + 1;
}

// cvvdp_metric.py:1057-1092; irfft(n=N) + fftshift restated as a direct inverse real DFT (N is odd).
int temporal_filters(const cvvdp_b200_params &P, double fps, cvvdp_b200_plan_info *info) {
    const int N = (int)(ceil(0.250 * fps / 2.0) * 2.0) + 1;
    if (N > CVVDP_MAX_FILTER_LEN) return -1;
    const int No = N / 2 + 1;
    const double pi = 3.14159265358979323846;
    for (int c = 0; c < 4; ++c) {
        std::vector<double> R(No);
        for (int k = 0; k < No; ++k) {
            const double om = (double)(float)((fps / 2.0) * k / (double)(No - 1 > 0 ? No - 1 : 1));
            if (c < 3) R[k] = exp(-pow(om, (double)P.beta_tf[c]) / (double)P.sigma_tf[c]);
            else {
                const double d = pow(om, (double)P.beta_tf[3]) - pow(5.0, (double)P.beta_tf[3]);
                R[k] = exp(-(d * d) / (double)P.sigma_tf[3]);
            }
        }
        for (int t = 0; t <= N / 2; ++t) {  // x[t] == x[N-t]: evaluate one half, mirror it (exactly symmetric taps)
            double x = R[0];
            for (int k = 1; k < No; ++k) x += 2.0 * R[k] * cos(2.0 * pi * k * t / (double)N);
            x /= (double)N;
            info->filters[c][(t + N / 2) % N] = (float)x;  // fftshift
            info->filters[c][(N - t + N / 2) % N] = (float)x;
        }
    }
    return N;
}

// csf.py:38-46 + interp.py:152-178 in fp32
void csf_row(const cvvdp_b200_csf_lut &lut, float rho, int ch, float *row) {
    float log_rho[CVVDP_CSF_LUT_N];
    for (int i = 0; i < CVVDP_CSF_LUT_N; ++i) log_rho[i] = log10f(lut.rho[i]);
    const float x = log10f(rho);
    int idx = 0;
    while (idx < CVVDP_CSF_LUT_N && log_rho[idx] < x) ++idx;  // searchsorted(side='left')
    idx = std::max(0, std::min(idx - 1, CVVDP_CSF_LUT_N - 2));
    const float x0 = log_rho[idx], x1 = log_rho[idx + 1];
    for (int l = 0; l < CVVDP_CSF_LUT_N; ++l) {
        const float y0 = lut.logS[ch][l][idx], y1 = lut.logS[ch][l][idx + 1];
        const float slope = (y1 - y0) / (x1 - x0);
        row[l] = y0 + slope * (x - x0);
    }
}

void to_display_dev(const cvvdp_b200_display &d, DisplayDev *o, int colorspace = CVVDP_CS_DKLD65) {
    o->eotf = d.eotf;
    o->gamma = d.gamma;
    o->Ypeak = d.Y_peak;
    const double Yblack = (double)d.Y_peak / (double)d.contrast;                 // display_model.py:374
    const double Yrefl = (double)d.E_ambient / 3.14159265358979323846 * d.k_refl;  // l.373
    o->Yblack = (float)Yblack;
    o->Yrefl = (float)Yrefl;
    o->exposure = d.exposure;
    o->lin_lo = (float)std::max(0.005, Yblack);
    // display_model.py:17-25, 255-256: (LMS2006_to_DKLd65 @ XYZ_to_LMS2006) @ rgb2xyz in fp32
    const float A[9] = {1.f, 1.f, 0.f, 1.f, -2.311130179947035f, 0.f, -1.f, -1.f, 50.977571328718781f};
    const float Bm[9] = {0.187596268556126f, 0.585168649077728f, -0.026384263306304f,
                         -0.133397430663221f, 0.405505777260049f, 0.034502127690364f,
                         0.000244379021663f, -0.000542995890619f, 0.019406849066323f};
    float AB[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float acc = 0.f;
            for (int k = 0; k < 3; ++k) acc = acc + A[i * 3 + k] * Bm[k * 3 + j];
            AB[i * 3 + j] = acc;
        }
    const float I3[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    const float *lhs = colorspace == CVVDP_CS_DKLD65 ? AB : (colorspace == CVVDP_CS_LMS2006 ? Bm : I3);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            if (colorspace == CVVDP_CS_RGB_LINEAR) {
                o->M[i * 3 + j] = I3[i * 3 + j];
                continue;
            }
            if (colorspace == CVVDP_CS_XYZ) {
                o->M[i * 3 + j] = d.rgb2xyz[i * 3 + j];
                continue;
            }
            float acc = 0.f;
            for (int k = 0; k < 3; ++k) acc = acc + lhs[i * 3 + k] * d.rgb2xyz[k * 3 + j];
            o->M[i * 3 + j] = acc;
        }
}

// video_source_yuv.py:196-207: limited-range weights/offsets for luma and chroma
void fill_yuv(const cvvdp_b200_yuv &y, int W, int H, YuvDev *o) {
    memset(o, 0, sizeof(*o));
    o->chroma = y.chroma;
    if (y.chroma == 0) return;
    o->W = W;
    o->H = H;
    const double sc = pow(2.0, (double)(y.bit_depth - 8));
    o->yw = (float)(1.0 / (sc * 219.0));
    o->yo = (float)(16.0 / 219.0);
    o->cw = (float)(1.0 / (sc * 224.0));
    o->co = (float)(128.0 / 224.0);
    o->m_rv = y.coef[0];
    o->m_gu = y.coef[1];
    o->m_gv = y.coef[2];
    o->m_bu = y.coef[3];
}

// elements of one planar YUV frame
long long yuv_frame_elems(const cvvdp_b200_yuv &y, int W, int H) {
    const long long ypix = (long long)W * H;
    return y.chroma == 444 ? 3 * ypix : (y.chroma == 422 ? 2 * ypix : ypix * 3 / 2);
}

size_t dtype_size(int dtype) {
    switch (dtype) {
        case CVVDP_DTYPE_U8: return 1;
        case CVVDP_DTYPE_U16:
        case CVVDP_DTYPE_F16: return 2;
        default: return 4;
    }
}

// frames (clip indices) the temporal stage reads to produce outputs [f0, f1)
void needed_frames(const cvvdp_b200_ctx *ctx, int f0, int f1, int *lo, int *hi) {
    const int fl = ctx->info.filter_len, F = ctx->job.n_frames;
    int mn = f1 - 1, mx = f1 - 1;
    for (int t = f0 - (fl - 1); t < f1; ++t) {
        int s = t;
        if (s < 0) {
            if (ctx->job.padding == CVVDP_PAD_REPLICATE) s = 0;
            else {
                const int m = F - 1, a = -s - 1;
                if (((a / m) & 1) == 0) s = (a % m) + 1;
                else {
                    s = t % m;
                    if (s < 0) s += m;
                }
            }
        }
        mn = std::min(mn, s);
        mx = std::max(mx, s);
    }
    *lo = mn;
    *hi = mx + 1;
}

int check_clip(cvvdp_b200_ctx *ctx, const cvvdp_b200_clip *c, int lo, int hi, const char *name, bool from_file = false) {
    if (!c || (!c->data && !from_file)) return fail(ctx, CVVDP_ERR_INVALID, "%s clip is null", name);
    if (lo < c->frame0 || hi > c->frame0 + c->n_frames)
        return fail(ctx, CVVDP_ERR_INVALID, "%s view holds frames [%d,%d) but frames [%d,%d) are needed", name,
                    c->frame0, c->frame0 + c->n_frames, lo, hi);
    return CVVDP_OK;
}

ClipView to_view(const cvvdp_b200_clip *c) {
    ClipView v;
    v.data = c->data;
    for (int i = 0; i < 5; ++i) v.s[i] = c->stride[i];
    v.frame0 = c->frame0;
    v.n_frames = c->n_frames;
    v.ring = 0;
    return v;
}

// ---- band kernel dispatch ------------------------------------------------------------------------
// The level-dependent flags (blur, heat map, beta == 2) and the feature mode select the instantiation.
template <int EW, bool CF, bool LAG, bool FEAT>
void launch_band_v(const BandArgs &ba, dim3 grid, cudaStream_t st, int variant) {
    typedef void (*BandFn)(const BandArgs);
    // (feature mode never comes with a heat map -- the plan refuses the pair, like the reference's ML metrics -- so those
    // four instantiations do not exist)
    static const BandFn table[8] = {
        k_band2<EW, CF, LAG, false, false, false, FEAT>, k_band2<EW, CF, LAG, false, false, true, FEAT>,
        FEAT ? nullptr : k_band2<EW, CF, LAG, false, !FEAT, false, false>, FEAT ? nullptr : k_band2<EW, CF, LAG, false, !FEAT, true, false>,
        k_band2<EW, CF, LAG, true, false, false, FEAT>,  k_band2<EW, CF, LAG, true, false, true, FEAT>,
        FEAT ? nullptr : k_band2<EW, CF, LAG, true, !FEAT, false, false>,  FEAT ? nullptr : k_band2<EW, CF, LAG, true, !FEAT, true, false>};
    static bool attr_set[8] = {false, false, false, false, false, false, false, false};
    typedef Band2Smem<EW, B2Lag<LAG>::DFR> Smem;
    BandFn kfn = table[variant];
    if (!attr_set[variant]) {  // one context per device and process in practice; the attribute is per function
        cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
        attr_set[variant] = true;
    }
    CVVDP_LAUNCH(kfn, grid, dim3(CVVDP_B2_THREADS), sizeof(Smem), st, ba);
}

void launch_band(const BandArgs &ba, dim3 grid, cudaStream_t st, int variant, bool feat) {
    if (feat) launch_band_v<CVVDP_BAND_EW, CVVDP_BAND_CF, CVVDP_BAND_LAG, true>(ba, grid, st, variant);
    else launch_band_v<CVVDP_BAND_EW, CVVDP_BAND_CF, CVVDP_BAND_LAG, false>(ba, grid, st, variant);
}

// One block of frames [f0, f1) (f1 - f0 <= block_frames), inputs resident on the device.
int run_block(cvvdp_b200_ctx *ctx, const cvvdp_b200_clip *test, const cvvdp_b200_clip *ref, int f0, int f1,
              float *q_dev, void *hm_dev, cudaStream_t st, int ring = 0) {
    const cvvdp_b200_job &job = ctx->job;
    const cvvdp_b200_plan_info &info = ctx->info;
    const cvvdp_b200_params &P = ctx->P;
    const int n = f1 - f0, L = info.n_bands, B = job.batch;
    const int pairs = B * n;
    const bool do_hm = job.heatmap != CVVDP_HEATMAP_NONE;
    const bool do_feat = job.features != 0 && ctx->feat_out != nullptr;
    const float eps = 1e-5f;

    // ---- temporal stage ----
    if (job.prefiltered) {  // cvvdp_metric.py:470-488: the source's four temporal channels are level 0
        PackArgs pk;
        pk.clip[0] = to_view(test);
        pk.clip[1] = to_view(ref);
        pk.clip[0].ring = pk.clip[1].ring = ring;
        pk.B = B;
        pk.H = job.height;
        pk.W = job.width;
        pk.f0 = f0;
        pk.f1 = f1;
        pk.out = ctx->lv[0].g;
        const long long npix = (long long)job.height * job.width;
        LaunchScope ls(ctx, st, CVVDP_K_TEMPORAL, 0, (double)npix * pairs * 2 * 32.0);
        auto kfn = k_pack_level0;
        CVVDP_LAUNCH(kfn, dim3((unsigned)((npix + 255) / 256), (unsigned)(pairs * 2)), dim3(256), 0, st, pk);
    } else {
        TemporalArgs ta;
        memset(&ta, 0, sizeof(ta));
        ta.clip[0] = to_view(test);
        ta.clip[1] = to_view(ref);
        ta.clip[0].ring = ta.clip[1].ring = ring;
        // frames that already are DKLd65 (plugin sources) pass through with an identity matrix
        to_display_dev(ctx->disp, &ta.dd, ctx->disp.eotf == CVVDP_EOTF_NONE ? CVVDP_CS_RGB_LINEAR : CVVDP_CS_DKLD65);
        ta.dtype = job.dtype;
        ta.cin = job.in_channels;
        ta.B = B;
        ta.H = job.height;
        ta.W = job.width;
        ta.F_total = job.n_frames;
        ta.f0 = f0;
        ta.f1 = f1;
        ta.fl = info.filter_len;
        ta.padding = job.padding;
        ta.out = ctx->lv[0].g;
        ta.flags = ctx->flags_dev;
        ta.mean0 = reinterpret_cast<float *>(ctx->flags_dev + 4);
        if (job.n_frames == 1) {  // image: R = DKL, no transient channel (cvvdp_metric.py:462-465)
            ta.taps[0][0] = ta.taps[1][0] = ta.taps[2][0] = 1.f;
            ta.taps[3][0] = 0.f;
        } else {
            for (int c = 0; c < 4; ++c)
                for (int k = 0; k < info.filter_len; ++k) ta.taps[c][k] = info.filters[c][info.filter_len - 1 - k];
        }
        const long long npix = (long long)job.height * job.width;
        dim3 grid((unsigned)((npix + CVVDP_TEMPORAL_THREADS - 1) / CVVDP_TEMPORAL_THREADS), (unsigned)(B * 2));
        const size_t smem = (size_t)info.filter_len * 3 * CVVDP_TEMPORAL_THREADS * sizeof(float);
        int wlo, whi;
        needed_frames(ctx, f0, f1, &wlo, &whi);
        const double bytes = (double)npix * B * 2 * ((double)(whi - wlo) * job.in_channels * dtype_size(job.dtype) + 16.0 * n);
        LaunchScope ls(ctx, st, CVVDP_K_TEMPORAL, 0, bytes);
        const int e = ctx->disp.eotf;
        const bool is_yuv = job.yuv.chroma != 0;
        fill_yuv(job.yuv, job.width, job.height, &ta.yuv);
        const bool use_lut = !is_yuv && job.dtype == CVVDP_DTYPE_U8 &&
                             (e == CVVDP_EOTF_SRGB || e == CVVDP_EOTF_PQ || e == CVVDP_EOTF_LINEAR || e == CVVDP_EOTF_GAMMA);
        // two-stage packed kernel: dense, 16-byte aligned planes made of whole 64-pixel warp segments
        // ... either planar (pixel stride 1) or channel-interleaved (HWC frames: pixel stride 3, channel stride 1), both videos alike
        bool dense = !is_yuv && npix % 64 == 0 && job.in_channels <= 3;
        const bool inter = dense && job.in_channels == 3 && ta.clip[0].s[4] == 3 && ta.clip[1].s[4] == 3;
        ta.inter = inter ? 1 : 0;
        for (int v = 0; v < 2 && dense; ++v) {
            const ClipView &cvw = ta.clip[v];
            const long long es = (long long)dtype_size(job.dtype);
            const long long sb = B > 1 ? cvw.s[0] : 0;  // (the stride of a singleton batch dimension is arbitrary)
            if (inter) {
                if (cvw.s[1] != 1 || cvw.s[3] != 3LL * job.width) dense = false;
                if (((uintptr_t)cvw.data) % 16 || (sb * es) % 16 || (cvw.s[2] * es) % 16) dense = false;
            } else {
                if (cvw.s[4] != 1 || cvw.s[3] != job.width) dense = false;
                if (((uintptr_t)cvw.data) % 16 || (sb * es) % 16 || (cvw.s[1] * es) % 16 || (cvw.s[2] * es) % 16) dense = false;
            }
        }
        // planar YUV: rows of whole 64-pixel segments (the two pixels of a thread share a row), frames anywhere
        bool yuv_2s = is_yuv && job.width % 64 == 0 && (job.dtype == CVVDP_DTYPE_U8 || job.dtype == CVVDP_DTYPE_U16);
        for (int v = 0; v < 2 && yuv_2s; ++v) {  // (the 16-byte pieces of the luma and chroma rows must be aligned)
            const ClipView &cvw = ta.clip[v];
            const long long es = (long long)dtype_size(job.dtype);
            if (((uintptr_t)cvw.data) % 16 || (cvw.s[0] * es) % 16 || (cvw.s[2] * es) % 16) yuv_2s = false;
        }
        if (yuv_2s) dense = true;
        const int t_esz = use_lut ? 1 : (int)dtype_size(job.dtype);
        const size_t smem_2s = t2s_smem_bytes(info.filter_len, t_esz, is_yuv ? 5 : 3);
        bool taps_symmetric = true;  // exact: the two-stage kernel adds mirrored frames before multiplying
        for (int c = 0; c < 4; ++c)
            for (int k = 0; k < info.filter_len / 2; ++k)
                if (ta.taps[c][k] != ta.taps[c][info.filter_len - 1 - k]) taps_symmetric = false;
        static const bool force_generic = getenv("CVVDP_B200_GENERIC_TEMPORAL") != nullptr;  // test hook: exercise the fallback
        // (one tap = images: the same staged, table-driven front end, the FIR degenerates to the centre tap)
        const bool two_stage = dense && info.filter_len >= 1 && info.filter_len <= 17 && taps_symmetric && !force_generic;
        dim3 grid_2s((unsigned)((npix / 64 + CVVDP_T2S_THREADS / 32 - 1) / (CVVDP_T2S_THREADS / 32)), (unsigned)(B * 2));
        bool launched = false;
#define CVVDP_TEMPORAL_CASE(FLV)                                                                  \
    case FLV: {                                                                                   \
        if (use_lut) {                                                                            \
            auto kfn = k_temporal_2s<FLV, CVVDP_T2S_LUT>;                                         \
            cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_2s); \
            CVVDP_LAUNCH(kfn, grid_2s, dim3(CVVDP_T2S_THREADS), smem_2s, st, ta);                 \
        } else if (is_yuv) {                                                                      \
            auto kfn = k_temporal_2s<FLV, CVVDP_T2S_YUV>;                                         \
            cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_2s); \
            CVVDP_LAUNCH(kfn, grid_2s, dim3(CVVDP_T2S_THREADS), smem_2s, st, ta);                 \
        } else {                                                                                  \
            auto kfn = k_temporal_2s<FLV, CVVDP_T2S_ANY>;                                         \
            cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_2s); \
            CVVDP_LAUNCH(kfn, grid_2s, dim3(CVVDP_T2S_THREADS), smem_2s, st, ta);                 \
        }                                                                                         \
        launched = true;                                                                          \
    } break;
        // images of 8-bit planes: the table front end alone (one frame per item in the block)
        if (two_stage && info.filter_len == 1 && use_lut && n == 1 && npix % 4 == 0) {
            const unsigned gx = (unsigned)std::min<long long>((npix + 1023) / 1024, (long long)ctx->num_sms * 8);
            auto kfn = k_image_lut;
            CVVDP_LAUNCH(kfn, dim3(gx, (unsigned)(B * 2)), dim3(256), 0, st, ta);
            launched = true;
        }
        if (two_stage && !launched) {
            switch (info.filter_len) {
#ifndef CVVDP_DEV_FAST  // development builds keep only the 30 and 60 fps specialisations
                CVVDP_TEMPORAL_CASE(3)
                CVVDP_TEMPORAL_CASE(5)
                CVVDP_TEMPORAL_CASE(7)
                CVVDP_TEMPORAL_CASE(11)
                CVVDP_TEMPORAL_CASE(13)
                CVVDP_TEMPORAL_CASE(15)
#endif
                CVVDP_TEMPORAL_CASE(1)
                CVVDP_TEMPORAL_CASE(9)
                CVVDP_TEMPORAL_CASE(17)
                default: break;
            }
        }
#undef CVVDP_TEMPORAL_CASE
        // long filters (frame rates above 64 fps) on dense planes: the shared-memory ring kernel
        const size_t smem_sr = tsr_smem_bytes(info.filter_len, use_lut);
        if (!launched && dense && !is_yuv && taps_symmetric && !force_generic && info.filter_len >= 19 &&
            smem_sr <= (size_t)std::min(ctx->max_smem_optin, 227 * 1024)) {
            dim3 grid_sr((unsigned)((npix / 64 + CVVDP_TSR_THREADS / 32 - 1) / (CVVDP_TSR_THREADS / 32)), (unsigned)(B * 2));
            if (use_lut) {
                auto kfn = k_temporal_sr<true>;
                cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sr);
                CVVDP_LAUNCH(kfn, grid_sr, dim3(CVVDP_TSR_THREADS), smem_sr, st, ta);
            } else {
                auto kfn = k_temporal_sr<false>;
                cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sr);
                CVVDP_LAUNCH(kfn, grid_sr, dim3(CVVDP_TSR_THREADS), smem_sr, st, ta);
            }
            launched = true;
        }
        if (!launched) {  // images, strided / permuted views, planar YUV of other widths, filters beyond 73 taps
            auto kfn = k_temporal;
            CVVDP_LAUNCH(kfn, grid, dim3(CVVDP_TEMPORAL_THREADS), smem, st, ta);
        }
    }
    // ---- Gaussian pyramid: level i -> i+1, unless the band kernel of level i computes level i+1 itself ----
    auto do_reduce = [&](int i) {
        ReduceArgs ra;
        ra.in = ctx->lv[i].g;
        ra.out = ctx->lv[i + 1].g;
        ra.h = ctx->lv[i].h;
        ra.w = ctx->lv[i].w;
        ra.hc = ctx->lv[i + 1].h;
        ra.wc = ctx->lv[i + 1].w;
        LaunchScope ls(ctx, st, CVVDP_K_REDUCE, i, (double)pairs * 2 * 16.0 * ((double)ra.h * ra.w + (double)ra.hc * ra.wc));
        static const bool no_tma_reduce = getenv("CVVDP_B200_NO_TMA") != nullptr;  // test hook: exercise the fallbacks
        if (ctx->lv[i].tm_ok && !no_tma_reduce) {  // persistent, TMA-staged, double-buffered
            Reduce2Args r2;
            r2.tm_in = ctx->lv[i].tm_reduce_in;
            r2.out = ra.out;
            r2.h = ra.h;
            r2.w = ra.w;
            r2.hc = ra.hc;
            r2.wc = ra.wc;
            r2.planes = pairs * 2;
            const long long tiles = (long long)((ra.wc + CVVDP_R2_TX - 1) / CVVDP_R2_TX) * ((ra.hc + 7) / 8) * r2.planes;
            const int grid2 = (int)std::min<long long>(tiles, (long long)ctx->num_sms * 4);
            auto kfn = k_reduce2<8>;
            CVVDP_LAUNCH(kfn, dim3(grid2), dim3(256), sizeof(Reduce2Smem<8>), st, r2);
        } else {
            dim3 grid((ra.wc + CVVDP_RTX - 1) / CVVDP_RTX, (ra.hc + CVVDP_RTY - 1) / CVVDP_RTY, pairs * 2);
            auto kfn = k_reduce;
            CVVDP_LAUNCH(kfn, grid, dim3(256), 0, st, ra);
        }
    };
    // ---- bands (each preceded by the reduce that produces its coarse level, when that is a separate launch) ----
    const bool is_image = job.n_frames == 1;
    const float ch_w[4] = {1.f, P.ch_chrom_w, P.ch_chrom_w, is_image ? 0.f : P.ch_trans_w};
    const float t_int = is_image ? P.image_int : 1.f;
    for (int i = 0; i + 1 < L; ++i) {
        static const bool no_tma = getenv("CVVDP_B200_NO_TMA") != nullptr;  // test hook: cp.async staging / plain reduce instead
        const bool fused_now = ctx->lv[i].fused && !no_tma;
        if (!fused_now) do_reduce(i);
        BandArgs ba;
        memset(&ba, 0, sizeof(ba));
        const LevelBuf &lv = ctx->lv[i];
        ba.fine = lv.g;
        ba.coarse = ctx->lv[i + 1].g;
        ba.lut = lv.lut;
        ba.partials = lv.partials;
        ba.hm = do_hm ? lv.hm : nullptr;
        ba.feat = do_feat ? lv.feat : nullptr;
        ba.feat_plane = (long long)pairs * lv.h * lv.w;
        {
            const float gain[4] = {1.f, 1.45f, 1.f, 1.f};  // the CSF rows carry the masking gain (plan), the features do not
            for (int c = 0; c < 4; ++c) ba.inv_gain[c] = 1.f / gain[c];
        }
        ba.h = lv.h;
        ba.w = lv.w;
        ba.hc = ctx->lv[i + 1].h;
        ba.wc = ctx->lv[i + 1].w;
        ba.do_blur = lv.do_blur;
        ba.mul = (i == 0) ? 1.f : 2.f;
        {
            const float x0 = log10f(ctx->lut.L_bkg[0]), x1 = log10f(ctx->lut.L_bkg[CVVDP_CSF_LUT_N - 1]);
            const double sc = (double)(CVVDP_CSF_LUT_N - 1) / ((double)x1 - (double)x0);
            ba.lut_a = (float)(log10(2.0) * sc);
            ba.lut_b = (float)(-(double)x0 * sc);
        }
        for (int k = 0; k < 2 * CVVDP_BHALO + 1; ++k) ba.kern[k] = ctx->blur_kern[k];
        ba.mc = powf(10.f, P.mask_c);
        for (int c = 0; c < 4; ++c) ba.q[c] = P.mask_q[c];
        ba.p = P.mask_p;
        for (int k = 0; k < 16; ++k) ba.X[k] = powf(2.f, P.xcm_weights[k]);
        ba.dmax = powf(10.f, P.d_max);
        ba.inv_dmax = 1.0f / ba.dmax;
        ba.eps = eps;
        ba.beta = P.beta;
        for (int c = 0; c < 4; ++c) ba.hm_w[c] = ch_w[c] * t_int;
        ba.hm_beta = P.beta_tch;
        ba.hm_scale = (i == 0) ? 1.f : 0.5f;
        ba.seg_rows = lv.seg_rows;
        ba.use_tma = (lv.tm_ok && ctx->lv[i + 1].tm_ok && !no_tma) ? 1 : 0;
        if (ba.use_tma) {
            ba.tm_fine = lv.tm;
            ba.tm_coarse = ctx->lv[i + 1].tm_as_coarse;
        }
        dim3 grid(lv.tiles_x, lv.tiles_y, pairs);
        LaunchScope ls(ctx, st, CVVDP_K_BAND, i,
                       (double)pairs * 2 * 16.0 * ((double)ba.h * ba.w + (double)ba.hc * ba.wc) +
                           (do_hm ? (double)pairs * 4.0 * ba.h * ba.w : 0.0));
        const int variant = (ba.do_blur ? 4 : 0) | (ba.hm ? 2 : 0) | (ba.beta == 2.0f ? 1 : 0);
        if (fused_now) {  // band + reduce in one kernel: writes level i+1
            ba.coarse_out = ctx->lv[i + 1].g;
            ba.tm_fine_a = lv.tm_fa;
            ba.tm_fine_b = lv.tm_fb;
            typedef void (*BandFn)(const BandArgs);
            static const BandFn tf[4] = {k_band2f<false, false>, k_band2f<false, true>, k_band2f<true, false>, k_band2f<true, true>};
            static bool attr_f[4] = {false, false, false, false};
            const int vf = variant & 3;
            if (!attr_f[vf]) {
                cudaFuncSetAttribute(tf[vf], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BandFSmem));
                attr_f[vf] = true;
            }
            BandFn kfn = tf[vf];
            CVVDP_LAUNCH(kfn, grid, dim3(CVVDP_B2_THREADS), sizeof(BandFSmem), st, ba);
            continue;
        }
        launch_band(ba, grid, st, variant, do_feat);
    }
    {
        BasebandArgs bb;
        memset(&bb, 0, sizeof(bb));
        const LevelBuf &lv = ctx->lv[L - 1];
        bb.g = lv.g;
        bb.lut = lv.lut;
        bb.partials = lv.partials;
        bb.hm = do_hm ? lv.hm : nullptr;
        bb.feat = do_feat ? lv.feat : nullptr;
        bb.npix = lv.h * lv.w;
        const float x0 = log10f(ctx->lut.L_bkg[0]), x1 = log10f(ctx->lut.L_bkg[CVVDP_CSF_LUT_N - 1]);
        const double sc = (double)(CVVDP_CSF_LUT_N - 1) / ((double)x1 - (double)x0);
        bb.lut_a = (float)(log10(2.0) * sc);
        bb.lut_b = (float)(-(double)x0 * sc);
        bb.eps = eps;
        bb.beta = P.beta;
        for (int c = 0; c < 4; ++c) bb.hm_w[c] = ch_w[c] * t_int * P.baseband_weight[c];
        bb.hm_beta = P.beta_tch;
        auto kfn = k_baseband;
        LaunchScope ls(ctx, st, CVVDP_K_BASEBAND, L - 1, (double)pairs * 2 * 16.0 * bb.npix);
        CVVDP_LAUNCH(kfn, dim3(pairs), dim3(256), 0, st, bb);
    }
    // ---- spatial pooling epilogue -> Q_per_ch ----
    {
        FinalizeArgs fa;
        memset(&fa, 0, sizeof(fa));
        for (int i = 0; i < L; ++i) {
            fa.partials[i] = ctx->lv[i].partials;
            fa.ntiles[i] = (i == L - 1) ? 1 : ctx->lv[i].tiles_x * ctx->lv[i].tiles_y;
            fa.npix[i] = ctx->lv[i].h * ctx->lv[i].w;
        }
        fa.L = L;
        fa.C = info.n_channels;
        fa.B = B;
        fa.n = n;
        fa.f_off = f0;
        fa.F_total = job.n_frames;
        fa.beta = P.beta;
        fa.eps = eps;
        fa.Q = q_dev;
        const int warps = pairs * L;
        auto kfn = k_finalize;
        double pbytes = 0;
        for (int i = 0; i < L; ++i) pbytes += (double)pairs * fa.ntiles[i] * 16.0;
        LaunchScope ls(ctx, st, CVVDP_K_FINALIZE, 0, pbytes);
        CVVDP_LAUNCH(kfn, dim3((warps * 32 + 127) / 128), dim3(128), 0, st, fa);
    }
    // ---- feature pooling (cvvdp_ml_metric.py:78-106) ----
    if (do_feat) {
        for (int i = 0; i < L; ++i) {
            const LevelBuf &lv = ctx->lv[i];
            FeaturePoolArgs fp;
            fp.feat = lv.feat;
            fp.feat_plane = (long long)pairs * lv.h * lv.w;
            fp.out = ctx->feat_out + lv.feat_off;
            fp.h = lv.h;
            fp.w = lv.w;
            fp.ps = ctx->feature_size;
            fp.ph = lv.ph;
            fp.pw = lv.pw;
            fp.C = info.n_channels;
            fp.n = n;
            fp.f_off = f0;
            fp.F_total = job.n_frames;
            auto kfn = k_feature_pool;
            LaunchScope ls(ctx, st, CVVDP_K_FEATURES, i, (double)pairs * 3 * 16.0 * lv.h * lv.w);
            CVVDP_LAUNCH(kfn, dim3(lv.pw, lv.ph, pairs), dim3(128), 0, st, fp);
        }
    }
    // ---- heat map ----
    if (do_hm) {
        for (int i = L - 2; i >= 0; --i) {
            ExpandAddArgs ea;
            ea.coarse = ctx->lv[i + 1].hm;
            ea.fine = ctx->lv[i].hm;
            ea.h = ctx->lv[i].h;
            ea.w = ctx->lv[i].w;
            ea.hc = ctx->lv[i + 1].h;
            ea.wc = ctx->lv[i + 1].w;
            const long long npix = (long long)ea.h * ea.w;
            auto kfn = k_expand_add;
            LaunchScope ls(ctx, st, CVVDP_K_HEATMAP, i, (double)n * 4.0 * (2.0 * npix + (double)ea.hc * ea.wc));
            CVVDP_LAUNCH(kfn, dim3((unsigned)((npix + 255) / 256), n), dim3(256), 0, st, ea);
        }
        const long long npix0 = (long long)job.height * job.width;
        if (job.heatmap == CVVDP_HEATMAP_RAW) {
            HeatmapOutArgs ha;
            ha.img = ctx->lv[0].hm;
            ha.out = (unsigned short *)hm_dev;
            ha.npix = npix0;
            ha.f_off = f0;
            ha.jod_a = P.jod_a;
            ha.jod_exp = P.jod_exp;
            auto kfn = k_heatmap_out;
            LaunchScope ls(ctx, st, CVVDP_K_HEATMAP, -1, (double)n * 6.0 * ha.npix);
            CVVDP_LAUNCH(kfn, dim3((unsigned)((ha.npix + 255) / 256), n), dim3(256), 0, st, ha);
        } else {  // coloured map: tone curve of this block's context image, then colour (visualize_diff_map.py:23-106)
            HmToneArgs ta;
            ta.lv0 = ctx->lv[0].g;
            ta.npix = npix0;
            ta.n = n;
            ta.minmax = ctx->hm_tone_dev;
            ta.hist = reinterpret_cast<int *>(ctx->hm_tone_dev + 2);
            ta.curve = reinterpret_cast<float *>(ctx->hm_tone_dev + 2 + CVVDP_HM_BINS);
            ta.dr = 0.6f;
            static const unsigned init[2] = {0x7f800000u, 0u};
            CU_CHECK(ctx, cudaMemcpyAsync(ctx->hm_tone_dev, init, sizeof(init), cudaMemcpyHostToDevice, st));
            CU_CHECK(ctx, cudaMemsetAsync(ctx->hm_tone_dev + 2, 0, CVVDP_HM_BINS * 4, st));
            const unsigned gx = (unsigned)std::min<long long>((npix0 + 255) / 256, 4LL * ctx->num_sms);
            {
                LaunchScope ls(ctx, st, CVVDP_K_HEATMAP, -2, (double)n * 16.0 * npix0);
                auto k1 = k_hm_minmax;
                CVVDP_LAUNCH(k1, dim3(gx, n), dim3(256), 0, st, ta);
            }
            {
                LaunchScope ls(ctx, st, CVVDP_K_HEATMAP, -3, (double)n * 16.0 * npix0);
                auto k2 = k_hm_hist;
                CVVDP_LAUNCH(k2, dim3(gx, n), dim3(256), 0, st, ta);
            }
            {
                LaunchScope ls(ctx, st, CVVDP_K_HEATMAP, -4, 0.0);
                auto k3 = k_hm_curve;
                CVVDP_LAUNCH(k3, dim3(1), dim3(CVVDP_HM_BINS), 0, st, ta);
            }
            HmColourArgs ca;
            memset(&ca, 0, sizeof(ca));
            ca.img = ctx->lv[0].hm;
            ca.lv0 = ctx->lv[0].g;
            ca.minmax = ta.minmax;
            ca.curve = ta.curve;
            ca.out = (unsigned short *)hm_dev;
            ca.npix = npix0;
            ca.f_off = f0;
            ca.F_total = job.n_frames;
            ca.jod_a = P.jod_a;
            ca.jod_exp = P.jod_exp;
            ca.dr = 0.6f;
            static const float thr[5][3] = {{0.2f, 0.2f, 1.0f}, {0.2f, 1.0f, 1.0f}, {0.2f, 1.0f, 0.2f}, {1.0f, 1.0f, 0.2f}, {1.0f, 0.2f, 0.2f}};
            static const float sup[3][3] = {{0.2f, 1.0f, 1.0f}, {1.0f, 1.0f, 1.0f}, {1.0f, 1.0f, 0.2f}};
            const bool is_thr = job.heatmap == CVVDP_HEATMAP_THRESHOLD;
            ca.n_map = is_thr ? 5 : 3;
            for (int i = 0; i < ca.n_map; ++i) {
                const float *c = is_thr ? thr[i] : sup[i];
                ca.map_in[i] = is_thr ? (float)(0.25f * (float)i) * 0.1f : (float)(0.5f * (float)i) * 0.3f;  // l.67, 77
                const float lum = c[0] * 0.212656f + c[1] * 0.715158f + c[2] * 0.072186f;                 // l.98
                for (int k = 0; k < 3; ++k) ca.map_ch[i][k] = c[k] / (lum + 0.0001f);
            }
            LaunchScope ls(ctx, st, CVVDP_K_HEATMAP, -1, (double)n * (4.0 + 16.0 + 6.0) * npix0);
            auto k4 = k_hm_colour;
            CVVDP_LAUNCH(k4, dim3((unsigned)((npix0 + 255) / 256), n), dim3(256), 0, st, ca);
        }
    }
    CU_CHECK(ctx, cudaGetLastError());
    return CVVDP_OK;
}

int fill_pool_args(const cvvdp_b200_ctx *ctx, PoolArgs *pa, int B, int C, int F, int L) {
    const cvvdp_b200_params &P = ctx->P;
    memset(pa, 0, sizeof(*pa));
    pa->B = B;
    pa->C = C;
    pa->F = F;
    pa->L = L;
    const float w[4] = {1.f, P.ch_chrom_w, P.ch_chrom_w, P.ch_trans_w};
    for (int c = 0; c < 4; ++c) {
        pa->ch_w[c] = w[c];
        pa->bb_w[c] = P.baseband_weight[c];
    }
    pa->beta_sch = P.beta_sch;
    pa->beta_tch = P.beta_tch;
    pa->beta_t = P.beta_t;
    pa->image_int = P.image_int;
    pa->jod_a = P.jod_a;
    pa->jod_exp = P.jod_exp;
    pa->eps = 1e-5f;
    return CVVDP_OK;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int cvvdp_b200_abi_version(void) { return CVVDP_B200_ABI_VERSION; }

const char *cvvdp_b200_last_error(const cvvdp_b200_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int cvvdp_b200_create(const cvvdp_b200_params *params, const cvvdp_b200_csf_lut *lut, int device,
                      cvvdp_b200_ctx **out) {
    if (!params || !lut || !out) return fail(nullptr, CVVDP_ERR_INVALID, "null argument");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, CVVDP_ERR_CUDA, "no CUDA device available (%s): the B200 path has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= ndev) return fail(nullptr, CVVDP_ERR_INVALID, "device %d out of range", device);
    if (!(params->pu_dilate == 0.f || params->pu_dilate == 3.f))
        return fail(nullptr, CVVDP_ERR_UNSUPPORTED, "pu_dilate must be 0 or 3 (got %g)", params->pu_dilate);
    cvvdp_b200_ctx *ctx = new cvvdp_b200_ctx();
    ctx->device = device;
    ctx->P = *params;
    ctx->lut = *lut;
    DeviceGuard dev_guard(device);
    if (dev_guard.err != cudaSuccess) {
        delete ctx;
        return fail(nullptr, CVVDP_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    }
    cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
    // torchvision _get_gaussian_kernel1d(kernel_size = 4*sigma+1, sigma), cvvdp_metric.py:158
    ctx->blur_pad = (int)(params->pu_dilate * 2);
    if (ctx->blur_pad > 0) {
        const int ks = 2 * ctx->blur_pad + 1;
        float pdf[2 * CVVDP_BHALO + 1], sum = 0.f;
        for (int i = 0; i < ks; ++i) {
            const float x = -(float)ctx->blur_pad + (float)i;
            pdf[i] = expf(-0.5f * (x / params->pu_dilate) * (x / params->pu_dilate));
            sum += pdf[i];
        }
        for (int i = 0; i < ks; ++i) ctx->blur_kern[i] = pdf[i] / sum;
    }
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->work_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, CVVDP_ERR_CUDA, "stream creation failed");
    }
    for (auto &s : ctx->stage) {
        cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&s.consumed, cudaEventDisableTiming);
    }
    if (cudaMalloc(&ctx->hm_tone_dev, (2 + CVVDP_HM_BINS + 2 * CVVDP_HM_BINS + 2) * 4) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->hm_ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->hm_copied, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->dev_done, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, CVVDP_ERR_NOMEM, "cannot allocate the heat-map tone buffers");
    }
    if (cudaMalloc(&ctx->flags_dev, kFlagsBytes) != cudaSuccess || cudaMemset(ctx->flags_dev, 0, kFlagsBytes) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, CVVDP_ERR_NOMEM, "cannot allocate the validation flags");
    }
    auto kr2 = k_reduce2<8>;
    if (cudaFuncSetAttribute(kr2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Reduce2Smem<8>)) != cudaSuccess) {
        const cudaError_t e = cudaGetLastError();
        delete ctx;
        return fail(nullptr, CVVDP_ERR_CUDA, "cannot raise the shared-memory limit of the reduce kernel: %s", cudaGetErrorString(e));
    }
    auto kt = k_temporal;  // its ring grows with the filter length: allow everything the SM has, minus its static part
    {
        cudaFuncAttributes fa;
        int stat = 0;
        if (cudaFuncGetAttributes(&fa, kt) == cudaSuccess) stat = (int)fa.sharedSizeBytes;
        if (cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 std::min(ctx->max_smem_optin, 227 * 1024) - stat) != cudaSuccess) {
            const cudaError_t e = cudaGetLastError();
            delete ctx;
            return fail(nullptr, CVVDP_ERR_CUDA, "cannot raise the shared-memory limit of the temporal kernel: %s", cudaGetErrorString(e));
        }
    }
    *out = ctx;
    return CVVDP_OK;
}

void cvvdp_b200_destroy(cvvdp_b200_ctx *ctx) {
    if (!ctx) return;
    DeviceGuard dev_guard(ctx->device);
    cudaDeviceSynchronize();
    free_plan(ctx);
    if (ctx->q_dev) cudaFree(ctx->q_dev);
    if (ctx->hm_dev) cudaFree(ctx->hm_dev);
    if (ctx->flags_dev) cudaFree(ctx->flags_dev);
    if (ctx->hm_tone_dev) cudaFree(ctx->hm_tone_dev);
    for (int i = 0; i < cvvdp_b200_ctx::kPinSlots; ++i) {
        if (ctx->pin_buf[i]) cudaFreeHost(ctx->pin_buf[i]);
        if (ctx->pin_done[i]) cudaEventDestroy(ctx->pin_done[i]);
    }
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    if (ctx->hm_ready) cudaEventDestroy(ctx->hm_ready);
    if (ctx->hm_copied) cudaEventDestroy(ctx->hm_copied);
    if (ctx->dev_done) cudaEventDestroy(ctx->dev_done);
    for (auto &s : ctx->stage) {
        if (s.copied) cudaEventDestroy(s.copied);
        if (s.consumed) cudaEventDestroy(s.consumed);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->work_stream) cudaStreamDestroy(ctx->work_stream);
    delete ctx;
}

int cvvdp_b200_set_display(cvvdp_b200_ctx *ctx, const cvvdp_b200_display *d) {
    if (!ctx || !d) return fail(ctx, CVVDP_ERR_INVALID, "null argument");
    if (d->eotf < CVVDP_EOTF_SRGB || d->eotf > CVVDP_EOTF_NONE) return fail(ctx, CVVDP_ERR_INVALID, "unknown EOTF id %d", d->eotf);
    if (!(d->ppd > 0.f)) return fail(ctx, CVVDP_ERR_INVALID, "ppd must be positive");
    ctx->disp = *d;
    ctx->have_display = true;
    ctx->planned = false;  // cvvdp_metric.py:264 (self.lpyr = None)
    return CVVDP_OK;
}

int cvvdp_b200_plan(cvvdp_b200_ctx *ctx, const cvvdp_b200_job *job, cvvdp_b200_plan_info *info_out) {
    if (!ctx || !job) return fail(ctx, CVVDP_ERR_INVALID, "null argument");
    if (!ctx->have_display) return fail(ctx, CVVDP_ERR_STATE, "set_display must be called before plan");
    if (job->batch < 1 || job->height < 4 || job->width < 4 || job->n_frames < 1)
        return fail(ctx, CVVDP_ERR_INVALID, "bad job shape B=%d H=%d W=%d F=%d (H, W >= 4)", job->batch, job->height,
                    job->width, job->n_frames);
    if (job->prefiltered) {
        if (job->in_channels != 4 || job->dtype != CVVDP_DTYPE_F32 || job->n_frames < 2 || job->yuv.chroma != 0)
            return fail(ctx, CVVDP_ERR_INVALID, "pre-filtered clips: fp32, four channels, a video");
    } else if (job->in_channels != 1 && job->in_channels != 3)
        return fail(ctx, CVVDP_ERR_INVALID, "The content must have either 1 or 3 color channels.");
    if (job->n_frames > 1 && !(job->fps > 0.f))
        return fail(ctx, CVVDP_ERR_INVALID, "When passing video sequences, you must set frames_per_second parameter");
    if (job->dtype < CVVDP_DTYPE_U8 || job->dtype > CVVDP_DTYPE_F32) return fail(ctx, CVVDP_ERR_INVALID, "unknown dtype %d", job->dtype);
    if (job->padding != CVVDP_PAD_REPLICATE && job->padding != CVVDP_PAD_SYMMETRIC)
        return fail(ctx, CVVDP_ERR_INVALID, "Unknown padding method");
    if (job->heatmap < CVVDP_HEATMAP_NONE || job->heatmap > CVVDP_HEATMAP_SUPRATHRESHOLD)
        return fail(ctx, CVVDP_ERR_INVALID, "unknown heat-map mode %d", job->heatmap);
    if (job->heatmap != CVVDP_HEATMAP_NONE && job->batch > 1)
        return fail(ctx, CVVDP_ERR_INVALID, "Heatmaps not supported when batches are used");  // cvvdp_metric.py:311-312
    if (job->heatmap != CVVDP_HEATMAP_NONE && job->features)
        return fail(ctx, CVVDP_ERR_UNSUPPORTED, "Currently cvvdp-ml metrics do not produce heatmaps");  // cvvdp_ml_metric.py:120-121
    if (ctx->disp.eotf == CVVDP_EOTF_HLG && job->in_channels != 3)
        return fail(ctx, CVVDP_ERR_UNSUPPORTED, "HLG needs three colour channels");
    if (job->yuv.chroma != 0) {
        const cvvdp_b200_yuv &yv = job->yuv;
        if (yv.chroma != 420 && yv.chroma != 422 && yv.chroma != 444)
            return fail(ctx, CVVDP_ERR_UNSUPPORTED, "Unsupported chroma subsampling %d", yv.chroma);
        if (job->in_channels != 3 || job->batch != 1) return fail(ctx, CVVDP_ERR_INVALID, "YUV input: three channels, batch of one");
        if (yv.bit_depth < 8 || yv.bit_depth > 16 || (yv.bit_depth == 8) != (job->dtype == CVVDP_DTYPE_U8) ||
            (yv.bit_depth > 8 && job->dtype != CVVDP_DTYPE_U16))
            return fail(ctx, CVVDP_ERR_INVALID, "YUV input: dtype U8 for 8 bits, U16 for 9..16 bits");
        if ((yv.chroma != 444 && (job->width & 1)) || (yv.chroma == 420 && (job->height & 1)))
            return fail(ctx, CVVDP_ERR_INVALID, "subsampled chroma needs even luma dimensions");
    }
    DeviceGuard dev_guard(ctx->device);
    CU_CHECK(ctx, dev_guard.err);
    CU_CHECK(ctx, cudaDeviceSynchronize());
    free_plan(ctx);
    ctx->job = *job;
    cvvdp_b200_plan_info &info = ctx->info;
    memset(&info, 0, sizeof(info));
    const int L = band_setup(job->width, job->height, (double)ctx->disp.ppd, &info);
    info.n_channels = job->n_frames == 1 ? 3 : 4;
    if (job->n_frames == 1 || job->prefiltered) info.filter_len = 1;
    else {
        const int fl = temporal_filters(ctx->P, (double)job->fps, &info);
        if (fl < 0) return fail(ctx, CVVDP_ERR_UNSUPPORTED, "temporal filter longer than %d taps", CVVDP_MAX_FILTER_LEN);
        info.filter_len = fl;
    }
    if ((size_t)info.filter_len * 3 * CVVDP_TEMPORAL_THREADS * sizeof(float) > (size_t)std::min(ctx->max_smem_optin, 227 * 1024))
        return fail(ctx, CVVDP_ERR_UNSUPPORTED, "temporal filter of %d taps does not fit in shared memory", info.filter_len);

    // workspace per frame of a block (all batch items)
    const size_t B = (size_t)job->batch;
    const bool do_hm = job->heatmap != CVVDP_HEATMAP_NONE;
    size_t per_frame = 0;
    for (int i = 0; i < L; ++i) {
        const size_t npix = (size_t)info.band_height[i] * info.band_width[i];
        per_frame += align_up(B * 2 * npix * sizeof(float4), 256);
        const size_t tiles = (i == L - 1) ? 1
                                          : (size_t)((info.band_width[i] + 47) / 48) *
                                                (info.band_height[i] / 16 + 1);  // upper bound
        per_frame += align_up(B * tiles * 4 * sizeof(float), 256);
        if (do_hm) per_frame += align_up(npix * sizeof(float), 256);
        if (job->features) per_frame += align_up(3 * B * npix * sizeof(float4), 256);
    }
    size_t limit = job->workspace_limit_bytes > 0 ? (size_t)job->workspace_limit_bytes : (size_t)64 << 30;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) limit = std::min(limit, free_b / 2);
    int nb = (int)std::min<size_t>(std::max<size_t>(limit / std::max<size_t>(per_frame, 1), 1), 128);
    if (job->max_block_frames > 0) nb = std::min(nb, job->max_block_frames);
    nb = std::min(nb, job->n_frames);
    if ((long long)B * nb * 2 > 65535) nb = std::max(1, (int)(65535 / (B * 2)));
    info.block_frames = nb;

    // arena layout
    ctx->lv.resize(L);
    size_t off = 0;
    std::vector<size_t> off_g(L), off_p(L), off_h(L), off_l(L), off_f(L);
    for (int i = 0; i < L; ++i) {
        LevelBuf &lv = ctx->lv[i];
        lv.h = info.band_height[i];
        lv.w = info.band_width[i];
        lv.do_blur = (ctx->blur_pad > 0 && lv.h > ctx->blur_pad && lv.w > ctx->blur_pad) ? 1 : 0;  // cvvdp_metric.py:965
        {   // column strips of EW - 12 pixels; the rows are split into segments only when there are too few CTAs
            const int sw = CVVDP_BAND_EW - 2 * CVVDP_BHALO;
            lv.tiles_x = (lv.w + sw - 1) / sw;
            // the split depends on the level geometry only, never on the batch or block size, so that
            // the summation order (hence every bit of Q_per_ch) is independent of how frames are
            // partitioned into blocks, shards or ranks
            int nseg = std::min(std::max((256 + lv.tiles_x - 1) / lv.tiles_x, 1), std::max(1, lv.h / 64));
            lv.seg_rows = (((lv.h + nseg - 1) / nseg) + 7) / 8 * 8;
            lv.tiles_y = (lv.h + lv.seg_rows - 1) / lv.seg_rows;
        }
        const size_t npix = (size_t)lv.h * lv.w;
        off_g[i] = off;
        off += align_up(B * nb * 2 * npix * sizeof(float4), 256);
        const size_t tiles = (i == L - 1) ? 1 : (size_t)lv.tiles_x * lv.tiles_y;
        off_p[i] = off;
        off += align_up(B * nb * tiles * 4 * sizeof(float), 256);
        off_h[i] = off;
        if (do_hm) off += align_up((size_t)nb * npix * sizeof(float), 256);
        off_f[i] = off;
        if (job->features) off += align_up(3 * B * nb * npix * sizeof(float4), 256);
        off_l[i] = off;
        off += align_up(CVVDP_CSF_LUT_N * sizeof(float4), 256);
    }
    {   // feature tensors: [B][F][ph][pw][C][6] per band, band after band (cvvdp_ml_metric.py:349-352)
        ctx->feature_size = (int)ceil((double)ctx->disp.ppd);
        long long foff = 0;
        for (int i = 0; i < L; ++i) {
            LevelBuf &lv = ctx->lv[i];
            const int ps = std::max(ctx->feature_size, 1);
            lv.ph = (lv.h + ps - 1) / ps;
            lv.pw = (lv.w + ps - 1) / ps;
            lv.feat_off = foff;
            foff += (long long)B * job->n_frames * lv.ph * lv.pw * info.n_channels * 6;
        }
        ctx->feat_total = foff;
        ctx->feat_out = nullptr;  // a new plan invalidates the registered buffer
    }
    ctx->arena_bytes = off;
    info.workspace_bytes = (int64_t)off;
    if (cudaMalloc(&ctx->arena, off) != cudaSuccess) {
        ctx->arena = nullptr;
        cudaGetLastError();
        return fail(ctx, CVVDP_ERR_NOMEM, "cannot allocate %zu bytes of workspace", off);
    }
    // per-band CSF rows (csf.py:38-51 + cvvdp_metric.py:705-709), pre-scaled for exp2:
    //   S * gain = 2^(row * log2(10) + log2(10^(sens_corr/20) * gain));  gain [1,1.45,1,1] only for the
    //   masked bands (cvvdp_metric.py:835-837), not for the baseband (711-712).
    const double log2_10 = log2(10.0);
    const double sens = pow(10.0, (double)ctx->P.sensitivity_correction / 20.0);
    const double gain[4] = {1.0, 1.45, 1.0, 1.0};
    // L2 promotion of the tensor maps: 64 / 128 / 256 bytes and none were timed on one box (round 2): within 1.5 %
    const int reduce_promo = 128, band_promo = 128;
    // CVVDP_B200_FUSED_REDUCE=1: the band kernel of a level computes the next pyramid level itself (k_band2f: no reduce
    // launch, the level is read from HBM once -- the literal "one kernel per level" of the north star).  Parity-green
    // on hardware, but between 2 % faster and 5 % slower than the separate reduce + k_band2 pair depending on the box
    // (profiles/r02_ab_fused_*.txt), so the pair stays the default.
    const bool use_fused = getenv("CVVDP_B200_FUSED_REDUCE") != nullptr;
    for (int i = 0; i < L; ++i) {
        LevelBuf &lv = ctx->lv[i];
        char *base = (char *)ctx->arena;
        lv.g = (float4 *)(base + off_g[i]);
        lv.partials = (float *)(base + off_p[i]);
        lv.hm = do_hm ? (float *)(base + off_h[i]) : nullptr;
        lv.feat = job->features ? (float4 *)(base + off_f[i]) : nullptr;
        lv.lut = (float4 *)(base + off_l[i]);
        lv.tm_ok = make_tensor_map(&lv.tm, lv.g, lv.w, lv.h, (int)(B * nb * 2), CVVDP_BAND_EW, CVVDP_B2_RB, 2, band_promo) &&
                   make_tensor_map(&lv.tm_as_coarse, lv.g, lv.w, lv.h, (int)(B * nb * 2), CVVDP_BAND_EW / 2 + 2, CVVDP_B2_CR, 2, band_promo) &&
                   make_tensor_map(&lv.tm_reduce_in, lv.g, lv.w, lv.h, (int)(B * nb * 2), CVVDP_R2_IW, Reduce2Smem<8>::IH, 1, reduce_promo);
        lv.fused = false;
        if (use_fused && i + 1 < L && lv.tm_ok && lv.do_blur && !job->features &&
            make_tensor_map(&lv.tm_fa, lv.g, lv.w, lv.h, (int)(B * nb * 2), 64, CVVDP_BF_FR, 2, band_promo) &&
            make_tensor_map(&lv.tm_fb, lv.g, lv.w, lv.h, (int)(B * nb * 2), 4, CVVDP_BF_FR, 2, band_promo))
            lv.fused = true;
        float rows[4][CVVDP_CSF_LUT_N];
        for (int c = 0; c < 4; ++c) csf_row(ctx->lut, info.rho_band[i], c, rows[c]);
        float packed[CVVDP_CSF_LUT_N][4];
        for (int l = 0; l < CVVDP_CSF_LUT_N; ++l)
            for (int c = 0; c < 4; ++c) {
                const double g = (i == L - 1) ? 1.0 : gain[c];
                packed[l][c] = (float)((double)rows[c][l] * log2_10 + log2(sens * g));
            }
        CU_CHECK(ctx, cudaMemcpy(lv.lut, packed, sizeof(packed), cudaMemcpyHostToDevice));
    }
    ctx->planned = true;
    if (info_out) *info_out = info;
    return CVVDP_OK;
}

int cvvdp_b200_process_device(cvvdp_b200_ctx *ctx, const cvvdp_b200_clip *test, const cvvdp_b200_clip *ref,
                              int frame_begin, int frame_end, float *q_per_ch_dev, void *heatmap_dev, void *stream) {
    if (!ctx) return CVVDP_ERR_INVALID;
    if (!ctx->planned) return fail(ctx, CVVDP_ERR_STATE, "plan must be called before process");
    if (frame_begin < 0 || frame_end > ctx->job.n_frames || frame_begin >= frame_end)
        return fail(ctx, CVVDP_ERR_INVALID, "bad frame range [%d,%d)", frame_begin, frame_end);
    if (!q_per_ch_dev) return fail(ctx, CVVDP_ERR_INVALID, "q_per_ch_dev is null");
    if (ctx->job.heatmap != CVVDP_HEATMAP_NONE && !heatmap_dev) return fail(ctx, CVVDP_ERR_INVALID, "heatmap_dev is null");
    DeviceGuard dev_guard(ctx->device);
    CU_CHECK(ctx, dev_guard.err);
    int lo, hi;
    needed_frames(ctx, frame_begin, frame_end, &lo, &hi);
    int rc;
    if ((rc = check_clip(ctx, test, lo, hi, "test")) != CVVDP_OK) return rc;
    if ((rc = check_clip(ctx, ref, lo, hi, "reference")) != CVVDP_OK) return rc;
    const int nb = ctx->info.block_frames;
    for (int f0 = frame_begin; f0 < frame_end; f0 += nb) {
        const int f1 = std::min(f0 + nb, frame_end);
        if ((rc = run_block(ctx, test, ref, f0, f1, q_per_ch_dev, heatmap_dev, (cudaStream_t)stream)) != CVVDP_OK) return rc;
    }
    // the workspace is still in use on the caller's stream: later calls on the context's own streams order after this
    CU_CHECK(ctx, cudaEventRecord(ctx->dev_done, (cudaStream_t)stream));
    ctx->dev_pending = true;
    return CVVDP_OK;
}

// ---- host-buffer path ---------------------------------------------------------------------------
namespace {
// ---- host -> device upload ------------------------------------------------------------------------
// Pinned (or registered) host memory goes straight to the copy engine.  PAGEABLE memory -- what a caller who
// just loaded a clip into a numpy array passes to predict() -- would make cudaMemcpyAsync fall back to the
// driver's single-threaded staging (measured 11 GB/s on the bench box against 52 GB/s pinned); instead the
// library bounces it through three pinned 32 MiB slots that several host threads fill while the previous slot
// is on the wire.
bool is_pageable(const void *p) {
#ifdef CVVDP_EMU
    (void)p;
    return getenv("CVVDP_B200_FORCE_STAGING") != nullptr;  // test hook: exercise the bounce path on the mock device
#else
    if (getenv("CVVDP_B200_FORCE_STAGING")) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
#endif
}

// Where the host bytes of a clip live: memory (`ptr`) or a file (`fd` >= 0, `ptr` then is the byte offset).
struct HostSrc {
    const char *ptr;
    int fd;
};
// One part of a copy into a pinned slot; false on a short read.
bool copy_part(char *dst, const HostSrc &src, size_t off, size_t n) {
    if (src.fd < 0) {
        memcpy(dst, src.ptr + off, n);
        return true;
    }
    size_t done = 0;
    while (done < n) {  // the page cache hands the bytes over without mapping the file (no page-table population, no unmap)
        const ssize_t r = pread(src.fd, dst + done, n - done, (off_t)((uintptr_t)src.ptr + off + done));
        if (r <= 0) return false;
        done += (size_t)r;
    }
    return true;
}
// Persistent helper threads for the staging copies (a slot is 32 MB, i.e. well under a millisecond of copying: creating
// and joining threads per slot cost a fifth of that).  One pool per process; the caller works on the parts as well.
class CopyPool {
  public:
    explicit CopyPool(int helpers) {
        for (int i = 0; i < helpers; ++i) workers_.emplace_back([this]() { loop(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_work_.notify_all();
        for (auto &t : workers_) t.join();
    }
    void run(int parts, const std::function<void(int)> &fn) {
        std::unique_lock<std::mutex> lk(m_);
        fn_ = &fn;
        parts_ = parts;
        next_ = 0;
        pending_ = parts;
        ++gen_;
        cv_work_.notify_all();
        take_parts(lk);
        cv_done_.wait(lk, [this]() { return pending_ == 0; });
        fn_ = nullptr;
    }

  private:
    void take_parts(std::unique_lock<std::mutex> &lk) {  // called with the lock held
        while (fn_ != nullptr && next_ < parts_) {
            const int i = next_++;
            const std::function<void(int)> *fn = fn_;
            lk.unlock();
            (*fn)(i);
            lk.lock();
            if (--pending_ == 0) cv_done_.notify_all();
        }
    }
    void loop() {
        std::unique_lock<std::mutex> lk(m_);
        unsigned long long seen = 0;
        for (;;) {
            cv_work_.wait(lk, [&]() { return stop_ || gen_ != seen; });
            if (stop_) return;
            seen = gen_;
            take_parts(lk);
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_work_, cv_done_;
    const std::function<void(int)> *fn_ = nullptr;
    int parts_ = 0, next_ = 0, pending_ = 0;
    unsigned long long gen_ = 0;
    bool stop_ = false;
};

bool parallel_copy(char *dst, const HostSrc &src, size_t n, int threads) {
    // CVVDP_B200_FORCE_STAGING=pool (test hook): tiny parts, so that small test clips go through the pool as well
    const char *hook = getenv("CVVDP_B200_FORCE_STAGING");
    const bool tiny = hook != nullptr && strcmp(hook, "pool") == 0;
    const size_t part = tiny ? (size_t)16 << 10 : (size_t)1 << 20;  // 1 MB parts, handed out dynamically
    if (threads <= 1 || n < 4 * part) return copy_part(dst, src, 0, n);
    static CopyPool pool(threads - 1);
    static std::mutex one_at_a_time;  // contexts of several threads share the pool
    std::lock_guard<std::mutex> guard(one_at_a_time);
    const int parts = (int)((n + part - 1) / part);
    std::vector<char> ok((size_t)parts, 1);
    const std::function<void(int)> fn = [&](int i) {
        const size_t o = (size_t)i * part;
        ok[(size_t)i] = copy_part(dst + o, src, o, std::min(part, n - o)) ? 1 : 0;
    };
    pool.run(parts, fn);
    return std::all_of(ok.begin(), ok.end(), [](char c) { return c != 0; });
}

int upload(cvvdp_b200_ctx *ctx, void *dst, HostSrc src, size_t bytes, bool pageable, cudaStream_t st) {
    if (!pageable && src.fd < 0) {
        CU_CHECK(ctx, cudaMemcpyAsync(dst, src.ptr, bytes, cudaMemcpyHostToDevice, st));
        return CVVDP_OK;
    }
    const size_t slot_bytes = (size_t)32 << 20;
    if (ctx->pin_bytes < slot_bytes) {
        for (int i = 0; i < cvvdp_b200_ctx::kPinSlots; ++i) {
            if (cudaMallocHost(&ctx->pin_buf[i], slot_bytes) != cudaSuccess) {
                cudaGetLastError();
                return fail(ctx, CVVDP_ERR_NOMEM, "cannot allocate the pinned bounce buffers");
            }
            CU_CHECK(ctx, cudaEventCreateWithFlags(&ctx->pin_done[i], cudaEventDisableTiming));
            CU_CHECK(ctx, cudaEventRecord(ctx->pin_done[i], st));
        }
        ctx->pin_bytes = slot_bytes;
    }
    static const int threads = (int)std::min(8u, std::max(2u, std::thread::hardware_concurrency() / 2));
    for (size_t off = 0; off < bytes; off += slot_bytes) {
        const size_t len = std::min(slot_bytes, bytes - off);
        const int slot = ctx->pin_next;
        ctx->pin_next = (ctx->pin_next + 1) % cvvdp_b200_ctx::kPinSlots;
        CU_CHECK(ctx, cudaEventSynchronize(ctx->pin_done[slot]));  // the copy that last used this slot has left it
        HostSrc part = src;
        part.ptr += off;
        if (!parallel_copy((char *)ctx->pin_buf[slot], part, len, threads))
            return fail(ctx, CVVDP_ERR_INVALID, "short read from the clip file (fd %d)", src.fd);
        CU_CHECK(ctx, cudaMemcpyAsync((char *)dst + off, ctx->pin_buf[slot], len, cudaMemcpyHostToDevice, st));
        CU_CHECK(ctx, cudaEventRecord(ctx->pin_done[slot], st));
    }
    return CVVDP_OK;
}

struct HostLayout {
    // frames are copied per "outer index" (dims whose stride exceeds the frame stride) as contiguous spans
    int outer_dims[4];
    int n_outer = 0;
    long long extent[5];
    bool ok = false;
};

HostLayout analyse_layout(const cvvdp_b200_clip *c, const long long extent[5]) {
    HostLayout hl;
    for (int i = 0; i < 5; ++i) hl.extent[i] = extent[i];
    const long long sF = c->stride[2];
    if (sF <= 0) return hl;
    // inner dims (stride < sF, extent > 1) must tile [0, sF) densely
    std::vector<int> inner;
    for (int d = 0; d < 5; ++d) {
        if (d == 2 || extent[d] == 1) continue;
        if (c->stride[d] == 0) continue;  // broadcast dim: nothing to copy per index
        if (c->stride[d] < sF) inner.push_back(d);
        else hl.outer_dims[hl.n_outer++] = d;
    }
    std::sort(inner.begin(), inner.end(), [&](int a, int b) { return c->stride[a] < c->stride[b]; });
    long long expect = 1;
    for (int d : inner) {
        if (c->stride[d] != expect) return hl;
        expect *= extent[d];
    }
    if (expect != sF) return hl;
    hl.ok = true;
    return hl;
}
}  // namespace

namespace {
int process_host_impl(cvvdp_b200_ctx *ctx, const cvvdp_b200_clip *test, const cvvdp_b200_clip *ref, const int fds[2],
                      const long long file_off[2], int frame_begin, int frame_end, float *q_per_ch_host, void *heatmap_host);
}
int cvvdp_b200_process_host(cvvdp_b200_ctx *ctx, const cvvdp_b200_clip *test, const cvvdp_b200_clip *ref,
                            int frame_begin, int frame_end, float *q_per_ch_host, void *heatmap_host) {
    const int fds[2] = {-1, -1};
    const long long offs[2] = {0, 0};
    return process_host_impl(ctx, test, ref, fds, offs, frame_begin, frame_end, q_per_ch_host, heatmap_host);
}
int cvvdp_b200_process_files(cvvdp_b200_ctx *ctx, const cvvdp_b200_clip *test, const cvvdp_b200_clip *ref, int fd_test,
                             int fd_ref, int64_t offset_test, int64_t offset_ref, int frame_begin, int frame_end,
                             float *q_per_ch_host, void *heatmap_host) {
    if (!ctx) return CVVDP_ERR_INVALID;
    if (fd_test < 0 || fd_ref < 0 || offset_test < 0 || offset_ref < 0) return fail(ctx, CVVDP_ERR_INVALID, "bad file descriptor or offset");
    const int fds[2] = {fd_test, fd_ref};
    const long long offs[2] = {(long long)offset_test, (long long)offset_ref};
    return process_host_impl(ctx, test, ref, fds, offs, frame_begin, frame_end, q_per_ch_host, heatmap_host);
}
namespace {
int process_host_impl(cvvdp_b200_ctx *ctx, const cvvdp_b200_clip *test, const cvvdp_b200_clip *ref, const int fds[2],
                      const long long file_off[2], int frame_begin, int frame_end, float *q_per_ch_host, void *heatmap_host) {
    if (!ctx) return CVVDP_ERR_INVALID;
    if (!ctx->planned) return fail(ctx, CVVDP_ERR_STATE, "plan must be called before process");
    if (frame_begin < 0 || frame_end > ctx->job.n_frames || frame_begin >= frame_end)
        return fail(ctx, CVVDP_ERR_INVALID, "bad frame range [%d,%d)", frame_begin, frame_end);
    if (!q_per_ch_host) return fail(ctx, CVVDP_ERR_INVALID, "q_per_ch_host is null");
    const cvvdp_b200_job &job = ctx->job;
    const bool do_hm = job.heatmap != CVVDP_HEATMAP_NONE;
    if (do_hm && !heatmap_host) return fail(ctx, CVVDP_ERR_INVALID, "heatmap_host is null");
    DeviceGuard dev_guard(ctx->device);
    CU_CHECK(ctx, dev_guard.err);
    int lo, hi, rc;
    needed_frames(ctx, frame_begin, frame_end, &lo, &hi);
    if ((rc = check_clip(ctx, test, lo, hi, "test", fds[0] >= 0)) != CVVDP_OK) return rc;
    if ((rc = check_clip(ctx, ref, lo, hi, "reference", fds[1] >= 0)) != CVVDP_OK) return rc;

    const long long ext_t[5] = {test->stride[0] ? job.batch : 1, job.in_channels, 0, job.height, job.width};
    const long long ext_r[5] = {ref->stride[0] ? job.batch : 1, job.in_channels, 0, job.height, job.width};
    HostLayout hl[2] = {analyse_layout(test, ext_t), analyse_layout(ref, ext_r)};
    if (job.yuv.chroma != 0) {  // planar YUV: a frame is one dense run of stride[2] elements, batch of one
        for (int v = 0; v < 2; ++v) {
            hl[v] = HostLayout();
            for (int d = 0; d < 5; ++d) hl[v].extent[d] = 1;
            hl[v].ok = (v == 0 ? test : ref)->stride[2] >= yuv_frame_elems(job.yuv, job.width, job.height);
        }
    }
    if (!hl[0].ok || !hl[1].ok)
        return fail(ctx, CVVDP_ERR_UNSUPPORTED,
                    "host clips must store each frame densely (dims with a stride below the frame stride must tile it)");
    const cvvdp_b200_clip *clips[2] = {test, ref};
    const bool pageable[2] = {fds[0] >= 0 || is_pageable(test->data), fds[1] >= 0 || is_pageable(ref->data)};
    const size_t esz = dtype_size(job.dtype);
    // The upload is pipelined in chunks smaller than the device-resident block size so that compute
    // starts after the first few frames have arrived.  The staging area is a RING of frames (frame f at
    // slot f % ring): the fl-1 history frames of a chunk are simply still there, so every input byte
    // crosses PCIe exactly once and no device-to-device shuffling competes with the kernels.
    const int fl = ctx->info.filter_len;
    // coloured heat maps take their tone-curve statistics over a block of frames (like the reference): the
    // chunks then are exactly the plan's blocks, the same partition as process_device
    const bool hm_blocks = job.heatmap == CVVDP_HEATMAP_THRESHOLD || job.heatmap == CVVDP_HEATMAP_SUPRATHRESHOLD;
    const int nb = hm_blocks ? ctx->info.block_frames : std::min(ctx->info.block_frames, std::max(16, fl - 1));
    const int ring = 2 * nb + 2 * fl;  // chunk k-1 (being computed) + chunk k (being uploaded) + history/look-ahead

    // staging ring: [outer index][ring][frame]
    size_t need[2];
    long long n_outer_idx[2];
    for (int v = 0; v < 2; ++v) {
        n_outer_idx[v] = 1;
        for (int k = 0; k < hl[v].n_outer; ++k) n_outer_idx[v] *= hl[v].extent[hl[v].outer_dims[k]];
        need[v] = (size_t)n_outer_idx[v] * ring * (size_t)clips[v]->stride[2] * esz;
    }
    {
        Staging &s = ctx->stage[0];
        if (s.bytes < std::max(need[0], need[1])) {
            for (auto &b : s.buf) {
                if (b) cudaFree(b);
                b = nullptr;
            }
            s.bytes = std::max(need[0], need[1]);
            for (auto &b : s.buf)
                if (cudaMalloc(&b, s.bytes) != cudaSuccess) {
                    b = nullptr;
                    s.bytes = 0;
                    cudaGetLastError();
                    return fail(ctx, CVVDP_ERR_NOMEM, "cannot allocate staging buffers");
                }
        }
    }
    const size_t q_bytes = (size_t)job.batch * ctx->info.n_channels * job.n_frames * ctx->info.n_bands * sizeof(float);
    if (ctx->q_dev_bytes < q_bytes) {
        if (ctx->q_dev) cudaFree(ctx->q_dev);
        CU_CHECK(ctx, cudaMalloc(&ctx->q_dev, q_bytes));
        ctx->q_dev_bytes = q_bytes;
    }
    if (ctx->dev_pending) {  // an unsynchronised process_device call may still be using the workspace
        CU_CHECK(ctx, cudaStreamWaitEvent(ctx->work_stream, ctx->dev_done, 0));
        CU_CHECK(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->dev_done, 0));
        ctx->dev_pending = false;
    }
    CU_CHECK(ctx, cudaMemsetAsync(ctx->q_dev, 0, q_bytes, ctx->work_stream));
    const int hm_ch = job.heatmap == CVVDP_HEATMAP_RAW ? 1 : 3;
    const size_t hm_bytes = do_hm ? (size_t)hm_ch * job.n_frames * job.height * job.width * 2 : 0;
    if (do_hm && ctx->hm_dev_bytes < hm_bytes) {
        if (ctx->hm_dev) cudaFree(ctx->hm_dev);
        CU_CHECK(ctx, cudaMalloc(&ctx->hm_dev, hm_bytes));
        ctx->hm_dev_bytes = hm_bytes;
    }

    // device views of the two rings: inner dims + frame stride as on the host, outer dims compacted
    cvvdp_b200_clip dev_clip[2];
    int order[2][4];
    long long dstr[2][5];
    for (int v = 0; v < 2; ++v) {
        const cvvdp_b200_clip *c = clips[v];
        for (int k = 0; k < hl[v].n_outer; ++k) order[v][k] = hl[v].outer_dims[k];
        std::sort(order[v], order[v] + hl[v].n_outer, [&](int a, int b) { return c->stride[a] < c->stride[b]; });
        for (int d = 0; d < 5; ++d) dstr[v][d] = c->stride[d];
        long long ostride = (long long)ring * c->stride[2];
        for (int k = 0; k < hl[v].n_outer; ++k) {
            dstr[v][order[v][k]] = ostride;
            ostride *= hl[v].extent[order[v][k]];
        }
        dev_clip[v] = *c;
        dev_clip[v].data = ctx->stage[0].buf[v];
        dev_clip[v].frame0 = 0;
        dev_clip[v].n_frames = ring;
        for (int d = 0; d < 5; ++d) dev_clip[v].stride[d] = dstr[v][d];
    }

    static const bool dbg_tl = getenv("CVVDP_B200_DEBUG_TIMELINE") != nullptr;
    static const bool skip_compute = getenv("CVVDP_B200_DEBUG_SKIP_COMPUTE") != nullptr;  // upload-only timing probe
    std::vector<cudaEvent_t> tl;  // per chunk: copy begin, copy end, compute begin, compute end
    auto tl_mark = [&](cudaStream_t s) {
        if (!dbg_tl) return;
        cudaEvent_t e;
        cudaEventCreateWithFlags(&e, 0);
        cudaEventRecord(e, s);
        tl.push_back(e);
    };
    int blk = 0, up_lo = 0, up_hi = 0;  // frames [up_lo, up_hi) are resident in the ring
    for (int f0 = frame_begin, f1; f0 < frame_end; f0 = f1, ++blk) {
        // full chunks, then a tapered tail (8, 4, 4 frames) so that little compute is left once the last
        // byte has arrived
        const int rem = frame_end - f0;
        f1 = f0 + (rem > nb || hm_blocks ? std::min(nb, rem) : (rem > 4 ? (rem + 1) / 2 : rem));
        Staging &sg = ctx->stage[blk & 1];  // events only; the data lives in the ring of stage[0]
        int wlo, whi;
        needed_frames(ctx, f0, f1, &wlo, &whi);
        if (whi - wlo > ring - nb) return fail(ctx, CVVDP_ERR_STATE, "internal: staging ring too small");
        // uploading this chunk overwrites slots last read by the chunk two steps back
        if (blk >= 2) CU_CHECK(ctx, cudaStreamWaitEvent(ctx->copy_stream, sg.consumed, 0));
        tl_mark(ctx->copy_stream);
        const int new_lo = (blk > 0 && wlo >= up_lo && wlo <= up_hi) ? std::min(up_hi, whi) : wlo;
        for (int v = 0; v < 2; ++v) {
            const cvvdp_b200_clip *c = clips[v];
            const long long sF = c->stride[2];
            long long idx[4] = {0, 0, 0, 0};
            for (long long oi = 0; oi < n_outer_idx[v]; ++oi) {
                long long hbase = 0, dbase = 0;
                for (int k = 0; k < hl[v].n_outer; ++k) {
                    hbase += idx[k] * c->stride[order[v][k]];
                    dbase += idx[k] * dstr[v][order[v][k]];
                }
                // frames [new_lo, whi) go to slots f % ring: at most two contiguous runs
                for (int fa = new_lo; fa < whi;) {
                    const int slot = fa % ring;
                    const int run = std::min(whi - fa, ring - slot);
                    HostSrc hs;  // a memory address, or the byte offset in the file
                    hs.fd = fds[v];
                    hs.ptr = (fds[v] >= 0 ? (const char *)(uintptr_t)file_off[v] : (const char *)c->data) +
                             (hbase + (long long)(fa - c->frame0) * sF) * esz;
                    if ((rc = upload(ctx, (char *)ctx->stage[0].buf[v] + (dbase + (long long)slot * sF) * esz, hs,
                                     (size_t)run * sF * esz, pageable[v], ctx->copy_stream)) != CVVDP_OK)
                        return rc;
                    fa += run;
                }
                for (int k = 0; k < hl[v].n_outer; ++k) {
                    if (++idx[k] < hl[v].extent[order[v][k]]) break;
                    idx[k] = 0;
                }
            }
        }
        up_lo = (new_lo == wlo) ? wlo : std::max(up_lo, whi - ring);
        up_hi = whi;
        CU_CHECK(ctx, cudaEventRecord(sg.copied, ctx->copy_stream));
        tl_mark(ctx->copy_stream);
        CU_CHECK(ctx, cudaStreamWaitEvent(ctx->work_stream, sg.copied, 0));
        tl_mark(ctx->work_stream);
        if (!skip_compute &&
            (rc = run_block(ctx, &dev_clip[0], &dev_clip[1], f0, f1, ctx->q_dev, ctx->hm_dev, ctx->work_stream, ring)) != CVVDP_OK)
            return rc;
        CU_CHECK(ctx, cudaEventRecord(sg.consumed, ctx->work_stream));
        tl_mark(ctx->work_stream);
        if (do_hm && !skip_compute) {  // this block's heat-map frames go home on their own stream while the next block runs
            const size_t fbytes = (size_t)job.height * job.width * 2;
            CU_CHECK(ctx, cudaEventRecord(ctx->hm_ready, ctx->work_stream));
            CU_CHECK(ctx, cudaStreamWaitEvent(ctx->d2h_stream, ctx->hm_ready, 0));
            for (int c = 0; c < hm_ch; ++c) {
                const size_t off = ((size_t)c * job.n_frames + f0) * fbytes;
                CU_CHECK(ctx, cudaMemcpyAsync((char *)heatmap_host + off, (char *)ctx->hm_dev + off, (size_t)(f1 - f0) * fbytes,
                                              cudaMemcpyDeviceToHost, ctx->d2h_stream));
            }
        }
    }
    CU_CHECK(ctx, cudaMemcpyAsync(q_per_ch_host, ctx->q_dev, q_bytes, cudaMemcpyDeviceToHost, ctx->work_stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->work_stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->copy_stream));
    if (do_hm) CU_CHECK(ctx, cudaStreamSynchronize(ctx->d2h_stream));
    if (dbg_tl) {
        for (size_t i = 0; i + 3 < tl.size(); i += 4) {
            float t[4];
            for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&t[k], tl[0], tl[i + k]);
            fprintf(stderr, "[cvvdp timeline] chunk %zu: copy %.2f-%.2f ms, compute %.2f-%.2f ms\n", i / 4, t[0], t[1], t[2], t[3]);
        }
        for (auto e : tl) cudaEventDestroy(e);
    }
    return CVVDP_OK;
}
}  // namespace

int cvvdp_b200_pool_device(cvvdp_b200_ctx *ctx, const float *q_dev, int B, int C, int F, int L, float *jod_dev, void *stream) {
    if (!ctx || !q_dev || !jod_dev) return fail(ctx, CVVDP_ERR_INVALID, "null argument");
    if (B < 1 || C < 1 || C > 4 || F < 1 || L < 1) return fail(ctx, CVVDP_ERR_INVALID, "bad Q_per_ch shape");
    DeviceGuard dev_guard(ctx->device);
    CU_CHECK(ctx, dev_guard.err);
    PoolArgs pa;
    fill_pool_args(ctx, &pa, B, C, F, L);
    pa.Q = q_dev;
    pa.jod = jod_dev;
    auto kfn = k_pool;
    {
        LaunchScope ls(ctx, (cudaStream_t)stream, CVVDP_K_POOL, 0, (double)B * C * F * L * 4.0);
        CVVDP_LAUNCH(kfn, dim3(B), dim3(256), 0, (cudaStream_t)stream, pa);
    }
    CU_CHECK(ctx, cudaGetLastError());
    return CVVDP_OK;
}

int cvvdp_b200_pool(cvvdp_b200_ctx *ctx, const float *q_host, int B, int C, int F, int L, float *jod_host) {
    if (!ctx || !q_host || !jod_host) return fail(ctx, CVVDP_ERR_INVALID, "null argument");
    if (B < 1 || C < 1 || C > 4 || F < 1 || L < 1) return fail(ctx, CVVDP_ERR_INVALID, "bad Q_per_ch shape");
    DeviceGuard dev_guard(ctx->device);
    CU_CHECK(ctx, dev_guard.err);
    const size_t qb = (size_t)B * C * F * L * sizeof(float);
    float *q_dev = nullptr, *j_dev = nullptr;
    CU_CHECK(ctx, cudaMalloc(&q_dev, qb + 256));
    if (cudaMalloc(&j_dev, B * sizeof(float)) != cudaSuccess) {
        cudaFree(q_dev);
        return fail(ctx, CVVDP_ERR_NOMEM, "cudaMalloc failed");
    }
    int rc = CVVDP_OK;
    if (cudaMemcpyAsync(q_dev, q_host, qb, cudaMemcpyHostToDevice, ctx->work_stream) != cudaSuccess) rc = CVVDP_ERR_CUDA;
    if (rc == CVVDP_OK) rc = cvvdp_b200_pool_device(ctx, q_dev, B, C, F, L, j_dev, ctx->work_stream);
    if (rc == CVVDP_OK && cudaMemcpyAsync(jod_host, j_dev, B * sizeof(float), cudaMemcpyDeviceToHost, ctx->work_stream) != cudaSuccess)
        rc = CVVDP_ERR_CUDA;
    if (cudaStreamSynchronize(ctx->work_stream) != cudaSuccess) rc = CVVDP_ERR_CUDA;
    cudaFree(q_dev);
    cudaFree(j_dev);
    if (rc == CVVDP_ERR_CUDA) return fail(ctx, rc, "CUDA error in pool: %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
}

int cvvdp_b200_frontend(cvvdp_b200_ctx *ctx, const cvvdp_b200_clip *src, int batch, int in_channels, int height, int width,
                        int dtype, int frame, int colorspace, float *dst_dev, int32_t *flags_dev, void *stream) {
    if (!ctx || !src || !src->data || !dst_dev) return fail(ctx, CVVDP_ERR_INVALID, "null argument");
    if (!ctx->have_display) return fail(ctx, CVVDP_ERR_STATE, "set_display must be called first");
    if (in_channels != 1 && in_channels != 3) return fail(ctx, CVVDP_ERR_INVALID, "The content must have either 1 or 3 color channels.");
    if (frame < src->frame0 || frame >= src->frame0 + src->n_frames) return fail(ctx, CVVDP_ERR_INVALID, "frame %d outside the view", frame);
    if (ctx->disp.eotf == CVVDP_EOTF_HLG && in_channels != 3) return fail(ctx, CVVDP_ERR_UNSUPPORTED, "HLG needs three colour channels");
    if (colorspace < CVVDP_CS_DKLD65 || colorspace > CVVDP_CS_LMS2006) return fail(ctx, CVVDP_ERR_INVALID, "unknown colour space id %d", colorspace);
    DeviceGuard dev_guard(ctx->device);
    CU_CHECK(ctx, dev_guard.err);
    FrontendArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.clip = to_view(src);
    to_display_dev(ctx->disp, &fa.dd, colorspace);
    fa.dtype = dtype;
    fa.cin = in_channels;
    fa.B = batch;
    fa.H = height;
    fa.W = width;
    fa.frame = frame - src->frame0;
    fa.dst = dst_dev;
    fa.flags = flags_dev;
    const long long npix = (long long)height * width;
    auto kfn = k_frontend;
    {
        LaunchScope ls(ctx, (cudaStream_t)stream, CVVDP_K_FRONTEND, 0, (double)npix * batch * in_channels * (dtype_size(dtype) + 4.0));
        CVVDP_LAUNCH(kfn, dim3((unsigned)((npix + 255) / 256), batch), dim3(256), 0, (cudaStream_t)stream, fa);
    }
    CU_CHECK(ctx, cudaGetLastError());
    return CVVDP_OK;
}

int cvvdp_b200_frontend_yuv(cvvdp_b200_ctx *ctx, const cvvdp_b200_clip *src, const cvvdp_b200_yuv *yuv, int batch, int height,
                            int width, int dtype, int frame, int colorspace, float *dst_dev, void *stream) {
    if (!ctx || !src || !src->data || !dst_dev || !yuv) return fail(ctx, CVVDP_ERR_INVALID, "null argument");
    if (!ctx->have_display) return fail(ctx, CVVDP_ERR_STATE, "set_display must be called first");
    if (yuv->chroma != 420 && yuv->chroma != 422 && yuv->chroma != 444)
        return fail(ctx, CVVDP_ERR_UNSUPPORTED, "Unsupported chroma subsampling %d", yuv->chroma);
    if (dtype != CVVDP_DTYPE_U8 && dtype != CVVDP_DTYPE_U16) return fail(ctx, CVVDP_ERR_INVALID, "YUV input: dtype U8 or U16");
    if ((yuv->chroma != 444 && (width & 1)) || (yuv->chroma == 420 && (height & 1)))
        return fail(ctx, CVVDP_ERR_INVALID, "subsampled chroma needs even luma dimensions");
    if (frame < src->frame0 || frame >= src->frame0 + src->n_frames) return fail(ctx, CVVDP_ERR_INVALID, "frame %d outside the view", frame);
    if (colorspace < CVVDP_CS_DKLD65 || colorspace > CVVDP_CS_LMS2006) return fail(ctx, CVVDP_ERR_INVALID, "unknown colour space id %d", colorspace);
    DeviceGuard dev_guard(ctx->device);
    CU_CHECK(ctx, dev_guard.err);
    FrontendArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.clip = to_view(src);
    to_display_dev(ctx->disp, &fa.dd, colorspace);
    fill_yuv(*yuv, width, height, &fa.yuv);
    fa.dtype = dtype;
    fa.cin = 3;
    fa.B = batch;
    fa.H = height;
    fa.W = width;
    fa.frame = frame - src->frame0;
    fa.dst = dst_dev;
    fa.flags = nullptr;
    const long long npix = (long long)height * width;
    auto kfn = k_frontend;
    {
        LaunchScope ls(ctx, (cudaStream_t)stream, CVVDP_K_FRONTEND, 0,
                       (double)batch * (yuv_frame_elems(*yuv, width, height) * dtype_size(dtype) + 12.0 * npix));
        CVVDP_LAUNCH(kfn, dim3((unsigned)((npix + 255) / 256), batch), dim3(256), 0, (cudaStream_t)stream, fa);
    }
    CU_CHECK(ctx, cudaGetLastError());
    return CVVDP_OK;
}

int cvvdp_b200_resize(cvvdp_b200_ctx *ctx, const float *src_dev, float *dst_dev, int channels, int height, int width,
                      int out_height, int out_width, int mode, int clip01, void *stream) {
    if (!ctx || !src_dev || !dst_dev) return fail(ctx, CVVDP_ERR_INVALID, "null argument");
    if (channels < 1 || height < 1 || width < 1 || out_height < 1 || out_width < 1)
        return fail(ctx, CVVDP_ERR_INVALID, "resize: empty plane %dx%dx%d -> %dx%d", channels, height, width, out_height, out_width);
    if (mode < CVVDP_RESIZE_NEAREST || mode > CVVDP_RESIZE_AREA) return fail(ctx, CVVDP_ERR_INVALID, "unknown resize mode %d", mode);
    DeviceGuard dev_guard(ctx->device);
    CU_CHECK(ctx, dev_guard.err);
    ResizeArgs ra;
    ra.src = src_dev;
    ra.dst = dst_dev;
    ra.C = channels;
    ra.H = height;
    ra.W = width;
    ra.OH = out_height;
    ra.OW = out_width;
    ra.mode = mode;
    ra.clip01 = clip01;
    const long long onp = (long long)out_height * out_width;
    auto kfn = k_resize;
    {
        LaunchScope ls(ctx, (cudaStream_t)stream, CVVDP_K_FRONTEND, 0, 4.0 * channels * ((double)height * width + (double)onp));
        CVVDP_LAUNCH(kfn, dim3((unsigned)((onp + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, ra);
    }
    CU_CHECK(ctx, cudaGetLastError());
    return CVVDP_OK;
}

int cvvdp_b200_input_stats(cvvdp_b200_ctx *ctx, cvvdp_b200_input_report *out, int reset) {
    if (!ctx || !out) return fail(ctx, CVVDP_ERR_INVALID, "null argument");
    DeviceGuard dev_guard(ctx->device);
    CU_CHECK(ctx, dev_guard.err);
    CU_CHECK(ctx, cudaDeviceSynchronize());  // the temporal kernels of every stream of this context have finished
    int h[4 + CVVDP_MEAN_SLOTS];
    CU_CHECK(ctx, cudaMemcpy(h, ctx->flags_dev, sizeof(h), cudaMemcpyDeviceToHost));
    out->out_of_range = h[0];
    out->nan = h[1];
    out->inf = h[2];
    double sum = 0.0;  // the kernels spread their partial sums over the slots (one address would serialise the atomics)
    for (int i = 0; i < CVVDP_MEAN_SLOTS; ++i) {
        float part;
        memcpy(&part, &h[4 + i], sizeof(float));
        sum += (double)part;
    }
    out->first_frame_sum = sum;
    if (reset) CU_CHECK(ctx, cudaMemset(ctx->flags_dev, 0, kFlagsBytes));
    return CVVDP_OK;
}

int64_t cvvdp_b200_launch_count(const cvvdp_b200_ctx *ctx) { return ctx ? ctx->launches : 0; }

int cvvdp_b200_temporal_filters(const cvvdp_b200_ctx *ctx, float fps, float *filters) {
    if (!ctx || !filters || !(fps > 0.f)) return CVVDP_ERR_INVALID;
    cvvdp_b200_plan_info tmp;
    memset(&tmp, 0, sizeof(tmp));
    const int n = temporal_filters(ctx->P, (double)fps, &tmp);
    if (n < 0) return CVVDP_ERR_UNSUPPORTED;
    memcpy(filters, tmp.filters, sizeof(tmp.filters));
    return n;
}

int cvvdp_b200_feature_layout(const cvvdp_b200_ctx *ctx, int band, int32_t *ph, int32_t *pw, int32_t *feature_size,
                              int64_t *float_offset) {
    if (!ctx || !ctx->planned) return CVVDP_ERR_STATE;
    const int L = (int)ctx->lv.size();
    if (band < 0 || band > L) return CVVDP_ERR_INVALID;
    if (feature_size) *feature_size = ctx->feature_size;
    if (band == L) {
        if (ph) *ph = 0;
        if (pw) *pw = 0;
        if (float_offset) *float_offset = ctx->feat_total;
        return CVVDP_OK;
    }
    if (ph) *ph = ctx->lv[band].ph;
    if (pw) *pw = ctx->lv[band].pw;
    if (float_offset) *float_offset = ctx->lv[band].feat_off;
    return CVVDP_OK;
}

int cvvdp_b200_set_feature_output(cvvdp_b200_ctx *ctx, float *features_dev) {
    if (!ctx || !ctx->planned) return CVVDP_ERR_STATE;
    if (features_dev && !ctx->job.features)
        return fail(ctx, CVVDP_ERR_STATE, "the current plan was made without job.features");
    ctx->feat_out = features_dev;
    return CVVDP_OK;
}

int cvvdp_b200_band_strip_width(const cvvdp_b200_ctx *ctx, int level) {
    if (!ctx || !ctx->planned || level < 0 || level + 1 >= (int)ctx->lv.size()) return 0;
    return CVVDP_BAND_EW - 2 * CVVDP_BHALO;
}

int cvvdp_b200_profile_enable(cvvdp_b200_ctx *ctx, int enable) {
    if (!ctx) return CVVDP_ERR_INVALID;
    ctx->prof = enable != 0;
    return CVVDP_OK;
}

int cvvdp_b200_profile_read(cvvdp_b200_ctx *ctx, cvvdp_b200_kernel_stat *out, int max_entries, int *n_entries) {
    if (!ctx || !out || !n_entries) return fail(ctx, CVVDP_ERR_INVALID, "null argument");
    DeviceGuard dev_guard(ctx->device);
    CU_CHECK(ctx, dev_guard.err);
    CU_CHECK(ctx, cudaDeviceSynchronize());
    int n = 0;
    for (auto &r : ctx->prof_recs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
        int k = 0;
        for (; k < n; ++k)
            if (out[k].kind == r.kind && out[k].level == r.level) break;
        if (k == n) {
            if (n >= max_entries) continue;
            out[n].kind = r.kind;
            out[n].level = r.level;
            out[n].launches = 0;
            out[n].total_ms = 0.f;
            out[n].algo_bytes = 0.0;
            ++n;
        }
        out[k].launches++;
        out[k].total_ms += ms;
        out[k].algo_bytes += r.bytes;
    }
    ctx->prof_recs.clear();
    *n_entries = n;
    return CVVDP_OK;
}

}  // extern "C"
