/*
 * cvvdp_b200.h -- C ABI of the B200-native ColorVideoVDP hot path (libcvvdp_b200.so).
 *
 * The reference (gfxdisp/ColorVideoVDP, pure Python/PyTorch) has no FFI; the entry points below are
 * what a binding for its hot path binds.  Each one names the reference interface it replaces
 * (paths relative to the reference tree).  Plain pointers and sizes only: no torch types, no C++
 * types, no exceptions across the boundary.  Every function returns CVVDP_OK (0) or a negative
 * error code; cvvdp_b200_last_error() gives the message.  One context = one CUDA device, not
 * re-entrant (same contract as one `pycvvdp.cvvdp` object, cvvdp_metric.py:108-140).
 *
 * Pointers named *_dev are device pointers on the context's device, *_host are host pointers.
 * `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 */
#ifndef CVVDP_B200_H
#define CVVDP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVVDP_B200_ABI_VERSION 6
#define CVVDP_MAX_BANDS 16
#define CVVDP_MAX_FILTER_LEN 129
#define CVVDP_CSF_LUT_N 32

enum { CVVDP_OK = 0, CVVDP_ERR_INVALID = -1, CVVDP_ERR_CUDA = -2, CVVDP_ERR_NOMEM = -3, CVVDP_ERR_STATE = -4,
       CVVDP_ERR_UNSUPPORTED = -5 };

/* Input element types accepted by video_source_array._get_frame (video_source.py:324-342). */
enum { CVVDP_DTYPE_U8 = 0,   /* /255 */
       CVVDP_DTYPE_U16 = 1,  /* uint16 (or int16 bit pattern, "& 0xFFFF"), /65535 */
       CVVDP_DTYPE_F16 = 2, CVVDP_DTYPE_F32 = 3 };

/* EOTFs of vvdp_display_photo_eotf.forward (display_model.py:333-365). */
enum { CVVDP_EOTF_SRGB = 0, CVVDP_EOTF_PQ = 1, CVVDP_EOTF_LINEAR = 2, CVVDP_EOTF_HLG = 3, CVVDP_EOTF_GAMMA = 4,
       CVVDP_EOTF_NONE = 5 /* input already is DKLd65 (frames from a third-party video_source plugin) */ };

/* Temporal padding (cvvdp_metric.py:506-532). */
enum { CVVDP_PAD_REPLICATE = 0, CVVDP_PAD_SYMMETRIC = 1 };

/* Target colour spaces of cvvdp_b200_frontend (linear_2_target_colorspace, display_model.py:241-276). */
enum { CVVDP_CS_DKLD65 = 0, CVVDP_CS_RGB_LINEAR = 1 /* forward() only */, CVVDP_CS_XYZ = 2, CVVDP_CS_LMS2006 = 3 };

/* Heat map (cvvdp_metric.py:117,396-401).  RAW: [1,1,F,H,W] fp16 difference map.  THRESHOLD / SUPRATHRESHOLD:
 * [1,3,F,H,W] fp16 colour map of visualize_diff_map (visualize_diff_map.py:48-106): colour look-up table times the
 * tone-mapped TEST sustained-achromatic context image (R[:,0], cvvdp_metric.py:399-401); like the reference, the
 * tone-curve statistics are taken over the block of frames processed in one pass (plan_info.block_frames). */
enum { CVVDP_HEATMAP_NONE = 0, CVVDP_HEATMAP_RAW = 1, CVVDP_HEATMAP_THRESHOLD = 2, CVVDP_HEATMAP_SUPRATHRESHOLD = 3 };
/* Interpolation of cvvdp_b200_resize: the `--full-screen-resize` choices of run_cvvdp.py:100. */
enum { CVVDP_RESIZE_NEAREST = 0, CVVDP_RESIZE_BILINEAR = 1, CVVDP_RESIZE_BICUBIC = 2, CVVDP_RESIZE_AREA = 3 };

/* cvvdp_parameters.json (cvvdp_metric.py:146-229).  Replaces cvvdp.load_config. */
typedef struct {
    float mask_p, mask_c;
    float mask_q[4];
    float xcm_weights[16];           /* log2 weights, 4x4 row-major [source ch][masked ch] (cvvdp_metric.py:758) */
    float beta, beta_t, beta_tch, beta_sch;
    float sensitivity_correction;    /* dB */
    float jod_a, jod_exp;
    float image_int;
    float ch_chrom_w, ch_trans_w;
    float baseband_weight[4];
    float d_max;                     /* log10 */
    float sigma_tf[4], beta_tf[4];
    float pu_dilate;                 /* sigma of the phase-uncertainty Gaussian; kernel = 4*sigma+1 taps */
} cvvdp_b200_params;

/* csf_lut_<version>.json (csf.py:8-23): log10-sensitivity tables [L_bkg][rho] for the four
 * perceptual channels A-sust (o0_c1), RG (o0_c2), YV (o0_c3), A-trans (o5_c1). */
typedef struct {
    float L_bkg[CVVDP_CSF_LUT_N];
    float rho[CVVDP_CSF_LUT_N];
    float logS[4][CVVDP_CSF_LUT_N][CVVDP_CSF_LUT_N];
} cvvdp_b200_csf_lut;

/* Display photometry + geometry (display_model.py:301-376, 503-526).  Replaces
 * vvdp_display_photo_eotf / vvdp_display_geometry as seen by cvvdp.set_display_model. */
typedef struct {
    int32_t eotf;            /* CVVDP_EOTF_* */
    float gamma;             /* CVVDP_EOTF_GAMMA: exponent; CVVDP_EOTF_HLG: system gamma */
    float Y_peak, contrast, E_ambient, k_refl, exposure;
    float rgb2xyz[9];        /* row-major RGB2X, RGB2Y, RGB2Z (color_spaces.json) */
    float ppd;               /* pixels per visual degree */
} cvvdp_b200_display;

/* Planar YUV input (raw .yuv frames / ffmpeg rawvideo), replaces YUVReader._fixed2float_upscale +
 * get_frame_rgb_tensor (video_source_yuv.py:153-233) and video_reader_yuv_pytorch.unpack
 * (video_source_file.py:261-324): limited-range unpack, bilinear chroma upsampling
 * (F.interpolate(scale_factor=2, mode='bilinear')), YCbCr->RGB matrix, clip to 0..1 -- fused into the
 * front end.  A frame is Y[h*w] U[ch*cw] V[ch*cw] (8-bit, or 16-bit little endian for bit_depth > 8);
 * the clip view's frame stride (stride[2], in elements) is the frame size, stride[0] the batch stride,
 * the other strides are ignored.  chroma == 0 means "not YUV" (RGB / grey planes as described by the clip). */
typedef struct {
    int32_t chroma;          /* 0, 420, 422 or 444 */
    int32_t bit_depth;       /* 8..16 */
    float coef[4];           /* R = Y + coef[0] Cr; G = Y + coef[1] Cb + coef[2] Cr; B = Y + coef[3] Cb */
} cvvdp_b200_yuv;

/* One prediction job (what cvvdp.predict_video_source derives from the video source, cvvdp_metric.py:304-355). */
typedef struct {
    int32_t batch;           /* B */
    int32_t height, width;   /* H, W */
    int32_t n_frames;        /* F of the whole clip (1 = image) */
    float fps;               /* 0 for images */
    int32_t in_channels;     /* 1 or 3 */
    int32_t dtype;           /* CVVDP_DTYPE_* */
    int32_t padding;         /* CVVDP_PAD_* */
    int32_t heatmap;         /* CVVDP_HEATMAP_* */
    int32_t max_block_frames;/* frames per pass (0 = choose from the workspace budget) */
    int64_t workspace_limit_bytes; /* 0 = default */
    cvvdp_b200_yuv yuv;      /* yuv.chroma != 0: the clips hold planar YUV frames (dtype U8 or U16, in_channels 3) */
    int32_t features;        /* != 0: also produce the per-band patch statistics of the ML heads (see cvvdp_b200_feature_layout) */
    int32_t prefiltered;     /* != 0: the clips hold the four temporal channels already ('DKLd65_trans' frames of a video
                              * source with is_temporally_filtered, cvvdp_metric.py:470-488): fp32, in_channels 4, display
                              * CVVDP_EOTF_NONE; the temporal filter is bypassed */
} cvvdp_b200_job;

typedef struct {
    int32_t n_bands;         /* L */
    int32_t n_channels;      /* C: 3 image, 4 video */
    int32_t filter_len;      /* fl (1 for images) */
    int32_t block_frames;    /* frames processed per pass */
    float rho_band[CVVDP_MAX_BANDS];       /* cpd; baseband already 0.1 (cvvdp_metric.py:685-686) */
    int32_t band_height[CVVDP_MAX_BANDS], band_width[CVVDP_MAX_BANDS];
    float filters[4][CVVDP_MAX_FILTER_LEN]; /* temporal filters F[c][0..fl-1] (cvvdp_metric.py:1057-1092) */
    int64_t workspace_bytes;
} cvvdp_b200_plan_info;

/* A strided view of a [B,C,F,H,W] clip: strides in ELEMENTS; batch stride 0 broadcasts a singleton
 * batch (video_source.py:247-252).  `frame0` is the clip frame index stored at F-index 0 of the view
 * and `n_frames` how many frames the view holds, so a caller may pass a window of a longer clip. */
typedef struct {
    const void *data;
    int64_t stride[5];       /* B, C, F, H, W */
    int32_t frame0, n_frames;
} cvvdp_b200_clip;

typedef struct cvvdp_b200_ctx cvvdp_b200_ctx;

/* ABI / build information: returns CVVDP_B200_ABI_VERSION. */
int cvvdp_b200_abi_version(void);
/* Message of the last error on this context (or the last create error when ctx is NULL). */
const char *cvvdp_b200_last_error(const cvvdp_b200_ctx *ctx);

/* cvvdp.__init__/load_config (cvvdp_metric.py:109-229) + castleCSF.__init__ (csf.py:8-25). */
int cvvdp_b200_create(const cvvdp_b200_params *params, const cvvdp_b200_csf_lut *lut, int device,
                      cvvdp_b200_ctx **out);
void cvvdp_b200_destroy(cvvdp_b200_ctx *ctx);

/* cvvdp.set_display_model (cvvdp_metric.py:246-264). */
int cvvdp_b200_set_display(cvvdp_b200_ctx *ctx, const cvvdp_b200_display *display);

/* The per-clip set-up of predict_video_source (cvvdp_metric.py:317-355): pyramid shape
 * (lpyr_dec.py:18-52), temporal filters (1057-1092), CSF rows per band (csf.py:38-46), block size
 * (565-594), workspace allocation. */
int cvvdp_b200_plan(cvvdp_b200_ctx *ctx, const cvvdp_b200_job *job, cvvdp_b200_plan_info *info);

/* The hot loop (cvvdp_metric.py:374-392 = read_block_of_frames 453-561 + process_block_of_frames
 * 660-751) for clip frames [frame_begin, frame_end) with test/reference views resident in HBM.
 * Writes Q_per_ch[b][c][f][band] for those frames into q_per_ch_dev (fp32, full [B,C,F,L] layout,
 * other frames untouched) and, when planned, the heat map ([1,1,F,H,W] raw or [1,3,F,H,W] coloured, fp16) into heatmap_dev.
 * Asynchronous on `stream`. */
int cvvdp_b200_process_device(cvvdp_b200_ctx *ctx, const cvvdp_b200_clip *test, const cvvdp_b200_clip *ref,
                              int frame_begin, int frame_end, float *q_per_ch_dev, void *heatmap_dev,
                              void *stream);

/* Same, with test/reference in HOST memory (pinned for full overlap): frames are uploaded in blocks
 * on a copy stream overlapped with compute, results are returned to host memory, and the call
 * returns after everything has completed.  q_per_ch_host: [B,C,F,L] fp32, written whole (zero outside the frame
 * range); heatmap_host: fp16 or NULL, only the frames of the range are written. */
int cvvdp_b200_process_host(cvvdp_b200_ctx *ctx, const cvvdp_b200_clip *test, const cvvdp_b200_clip *ref,
                            int frame_begin, int frame_end, float *q_per_ch_host, void *heatmap_host);

/* Same, with the frames read straight from two open files (raw planar .yuv files, video_source_yuv.py:77-144, or any
 * file of dense frames): `data` of the views is ignored; element i of the view of video v lives at byte
 * offset_v + i * sizeof(dtype) of fd_v, addressed through the strides like memory.  The upload threads pread() into the
 * pinned staging slots, so the file is neither mapped nor copied on the way.  The descriptors stay the caller's. */
int cvvdp_b200_process_files(cvvdp_b200_ctx *ctx, const cvvdp_b200_clip *test, const cvvdp_b200_clip *ref, int fd_test,
                             int fd_ref, int64_t offset_test, int64_t offset_ref, int frame_begin, int frame_end,
                             float *q_per_ch_host, void *heatmap_host);

/* do_pooling_and_jods + met2jod (cvvdp_metric.py:610-658) on a host Q_per_ch [B,C,F,L]; jod_host: [B]. */
int cvvdp_b200_pool(cvvdp_b200_ctx *ctx, const float *q_per_ch_host, int B, int C, int F, int L, float *jod_host);
/* Same on device buffers, asynchronous on `stream` (used right after process_device / the all-reduce). */
int cvvdp_b200_pool_device(cvvdp_b200_ctx *ctx, const float *q_per_ch_dev, int B, int C, int F, int L,
                           float *jod_dev, void *stream);

/* vvdp_display_photometry.source_2_target_colorspace(frame, colorspace) for one strided [B,C,1,H,W]
 * frame (display_model.py:206-276 + video_source.py:320-346): dst_dev is dense fp32 [B,3 or 1,H,W];
 * colorspace is a CVVDP_CS_* id (CVVDP_CS_RGB_LINEAR = vvdp_display_photo_eotf.forward alone).
 * Also counts values outside 0..1 / NaN / Inf (the warnings of display_model.py:335-337 and
 * video_source.py:48-59) into flags_dev[0..2] when not NULL. */
int cvvdp_b200_frontend(cvvdp_b200_ctx *ctx, const cvvdp_b200_clip *src, int batch, int in_channels, int height,
                        int width, int dtype, int frame, int colorspace, float *dst_dev, int32_t *flags_dev,
                        void *stream);
/* Same for a planar YUV frame; with the context's display set to CVVDP_EOTF_NONE and colorspace
 * CVVDP_CS_RGB_LINEAR the result is the display-encoded RGB of YUVReader.get_frame_rgb_tensor. */
int cvvdp_b200_frontend_yuv(cvvdp_b200_ctx *ctx, const cvvdp_b200_clip *src, const cvvdp_b200_yuv *yuv, int batch,
                            int height, int width, int dtype, int frame, int colorspace, float *dst_dev, void *stream);

/* Full-screen resize of one decoded frame: torch.nn.functional.interpolate(frame, size=(out_height, out_width),
 * mode=...) followed by clip(0, 1) when clip01 != 0, as video_source_yuv.py:257-260,333-336 and
 * video_source_file.py:280-287 apply it between the YCbCr matrix and the display model.  src_dev / dst_dev are
 * dense fp32 planes [channels][height][width] / [channels][out_height][out_width] on the context's device;
 * asynchronous on `stream`. */
int cvvdp_b200_resize(cvvdp_b200_ctx *ctx, const float *src_dev, float *dst_dev, int channels, int height, int width,
                      int out_height, int out_width, int mode, int clip01, void *stream);

/* Input validation of the fused path -- replaces the checks the reference makes frame by frame on the host:
 * vvdp_display_photo_eotf.forward "Pixel outside the valid range 0-1" (display_model.py:335-337, only for EOTFs that
 * clamp), video_source.check_if_valid NaN / Inf / first-frame mean (video_source.py:48-72).  The temporal kernels
 * count, per warp that saw one, floating-point input values outside 0..1, NaN and Inf (integer clips cannot carry
 * any), and accumulate the achromatic DKL channel of clip frame 0 of the TEST video over all batch items
 * (first_frame_sum; divide by batch * height * width; only meaningful when a process call started at frame 0).
 * cvvdp_b200_input_stats synchronises the device, reports the counters accumulated since the last reset and
 * optionally resets them.  A NaN input makes the reference fail with "Must not be nan" (cvvdp_metric.py:906-907);
 * the caller of this ABI is expected to do the same when nan != 0. */
typedef struct {
    int64_t out_of_range, nan, inf;
    double first_frame_sum;
} cvvdp_b200_input_report;
int cvvdp_b200_input_stats(cvvdp_b200_ctx *ctx, cvvdp_b200_input_report *out, int reset);

/* Per-kernel timing for bench.py's roofline: when enabled, every launch is bracketed by CUDA events on
 * its stream.  profile_read synchronises the device, aggregates by (kind, pyramid level) and resets.
 * algo_bytes is the ALGORITHMIC HBM traffic of those launches:
 * temporal = input bytes read once + 32 B/pixel written, reduce = 32 read + 8 written per input
 * pixel-pair, band = 32 + 8 read per level pixel-pair, ... */
enum { CVVDP_K_TEMPORAL = 0, CVVDP_K_REDUCE = 1, CVVDP_K_BAND = 2, CVVDP_K_BASEBAND = 3, CVVDP_K_FINALIZE = 4,
       CVVDP_K_HEATMAP = 5, CVVDP_K_POOL = 6, CVVDP_K_FRONTEND = 7, CVVDP_K_FEATURES = 8 };
typedef struct {
    int32_t kind, level, launches;
    float total_ms;
    double algo_bytes;
} cvvdp_b200_kernel_stat;
int cvvdp_b200_profile_enable(cvvdp_b200_ctx *ctx, int enable);
int cvvdp_b200_profile_read(cvvdp_b200_ctx *ctx, cvvdp_b200_kernel_stat *out, int max_entries, int *n_entries);

/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t cvvdp_b200_launch_count(const cvvdp_b200_ctx *ctx);

/* Temporal filters of the four channels at `fps` -- replaces cvvdp.get_temporal_filters (pycvvdp/cvvdp_metric.py:1057-1092):
 * writes F[c][0..n-1], c = A-sust, RG, YV, A-trans, to `filters` (4 rows of CVVDP_MAX_FILTER_LEN floats) and returns n
 * (odd), or a negative error code.  These are exactly the taps the temporal kernel uses. */
int cvvdp_b200_temporal_filters(const cvvdp_b200_ctx *ctx, float fps, float *filters);

/* Feature mode (job.features != 0) -- replaces cvvdp_ml_base.extract_features / cvvdp_feature_pooling
 * (pycvvdp/cvvdp_ml_metric.py:78-106, 206-298, 302-352): for every band, mean and variance of |T_f| S, |R_f| S and D
 * over feature_size x feature_size patches (feature_size = ceil(ppd), ragged border patches averaged over the
 * pixels they cover).  The tensor of band `band` is fp32 [B][F][ph][pw][C][6] with the last axis
 * (mean_T, var_T, mean_R, var_R, mean_D, var_D); all bands live in ONE device buffer, band after band.
 * cvvdp_b200_feature_layout reports ph, pw and the float offset of a band (offset of band n_bands = total
 * floats); cvvdp_b200_set_feature_output registers the device buffer that the next process_device /
 * process_host calls fill for the frames they evaluate (NULL detaches it). */
int cvvdp_b200_feature_layout(const cvvdp_b200_ctx *ctx, int band, int32_t *ph, int32_t *pw, int32_t *feature_size,
                              int64_t *float_offset);
int cvvdp_b200_set_feature_output(cvvdp_b200_ctx *ctx, float *features_dev);

/* Column-strip width of the band kernel for pyramid level `level` of the current plan (0: baseband or no plan).
 * Introspection for tests and profiling only. */
int cvvdp_b200_band_strip_width(const cvvdp_b200_ctx *ctx, int level);

#ifdef __cplusplus
}
#endif
#endif /* CVVDP_B200_H */
