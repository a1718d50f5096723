/*
 * ukbb_fcn.h -- C ABI of the B200-native FCN segmentation engine.
 *
 * Drop-in boundary for the device side of ukbb_cardiac's deploy path.  Each entry
 * point names the reference interface it replaces (paths relative to the
 * reference repository, baiwenjia/ukbb_cardiac):
 *
 *   ukbb_fcn_create        <- tf.train.import_meta_graph + saver.restore
 *                             (common/deploy_network.py:48-49): the graph is
 *                             build_FCN (common/network.py:170-230) with the
 *                             hyper-parameters of common/train_network.py:174-195;
 *                             weights arrive as host float32 arrays in TF layout.
 *   ukbb_fcn_forward       <- sess.run(['prob:0','pred:0'], feed_dict={'image:0': ...,
 *                             'training:0': False}) (deploy_network.py:110-111, 195-196)
 *                             fused with the transpose + crop of :114-116 / :199-200.
 *   ukbb_fcn_preprocess    <- rescale_intensity(image, (1, 99)) (common/image_utils.py:70-77,
 *                             called at deploy_network.py:89, 179) fused with the
 *                             pad-to-multiple-of-16 of deploy_network.py:97-100, 185-188
 *                             and the (X,Y,Z)->(Z,X,Y,1) float32 gather of :105-107.
 *   ukbb_fcn_segment_host  <- one whole iteration of the per-subject body
 *                             deploy_network.py:89-116 on HOST buffers (H2D, preprocess,
 *                             forward over all Z*T slices, D2H of the label volume).
 *   ukbb_fcn_class_counts  <- np.sum(pred == k, axis=(0,1,2)) of deploy_network.py:127-130
 *                             (per-slice class histogram emitted by the classifier).
 *   ukbb_cc_stats          <- measure.label + per-component np.sum of get_largest_cc / remove_small_cc
 *                             (common/image_utils.py:227-249) and skimage.measure.label of
 *                             atrium_pass_quality_control (common/cardiac_utils.py:1629-1640).
 *
 * Conventions
 *   - Plain C, no torch / C++ types.  Every function returns 0 on success and a
 *     negative UKBB_E_* code on failure; ukbb_last_error() returns a per-thread
 *     message.  The library never calls exit() and has NO CPU fallback: without a
 *     CUDA device every compute entry point fails with UKBB_E_CUDA.
 *   - A handle is bound to one device and is not re-entrant; distinct handles may
 *     be driven from distinct host threads.  All compute calls are asynchronous on
 *     the given stream (a cudaStream_t passed as void*; NULL = legacy default stream).
 *   - Memory layout: all image-like buffers are in NIfTI memory order, X fastest:
 *     volume [T][Z][Y][X]; padded network input [N=Z*T][Y2][X2]; labels [N][Y][X];
 *     logits / prob [N][Y2][X2][C].  TensorFlow's NHWC view of the same slice has
 *     H = X and W = Y; the library transposes the 3x3 kernels once at create time
 *     instead of transposing every image.
 */
#ifndef UKBB_FCN_H_
#define UKBB_FCN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UKBB_OK              0
#define UKBB_E_INVALID      -1   /* bad argument / shape */
#define UKBB_E_CUDA         -2   /* CUDA runtime or driver error (message has the detail) */
#define UKBB_E_NOMEM        -3
#define UKBB_E_UNSUPPORTED  -4   /* e.g. tensor-core mode on a non-sm_100 device */

/* arithmetic modes */
#define UKBB_MODE_FP32  0   /* FP32 CUDA-core kernels: the exactness mode */
#define UKBB_MODE_BF16  2   /* BF16 operands, FP32 accumulate on tcgen05 tensor cores */
#define UKBB_MODE_FP16  3   /* FP16 operands, FP32 accumulate: same kernels and tensor-core rate as
                               BF16, 11-bit significand (activations are post-BN/ReLU, O(1..100)) */
/* Split-operand tensor-core modes (the "3x" scheme, cf. 3xTF32): every activation and every weight is carried as
 * the sum of TWO 16-bit values, v = hi + lo with hi = rn16(v), lo = rn16(v - hi), and every product is evaluated as
 * hi.hi + lo.hi + hi.lo in the FP32 accumulator (three tcgen05.mma per K step; the lo.lo term is dropped).
 * Effective significand: 16 bits (BF16X3) / 22 bits (FP16X3, activations clamped to +-65504).  These are the
 * tensor-core modes that meet the >= 99.9 % label agreement / Dice >= 0.999 tolerance against the float32 reference
 * on random-init weights; plain BF16 / FP16 do not. */
#define UKBB_MODE_BF16X3  4
#define UKBB_MODE_FP16X3  5
/* The "2x" scheme: FP16 pieces as FP16X3, but BOTH correction products of a K step are ONE FP8 (E4M3) tensor-core instruction:
 *   a.w  ~=  a_hi.w_hi  +  2^-15 [e4m3(a_lo 2^11) | e4m3(a_hi)] . [e4m3(w_hi 2^4) | e4m3(w_lo 2^15)]
 * (K = 32 FP8 elements occupy the 32 bytes of an FP16 K = 16 step, so tensors keep the two-plane layout; the 2^-15 is the
 * scale-input-d of the first FP16 instruction).  Two instructions per K step instead of three; the corrections carry 4 significant
 * bits, so the product error is ~2^-16.7 relative (FP16: 2^-11.8, FP16X3: 2^-22): ~28x tighter than FP16 -- measured by
 * experiments/x2f8_probe.cu -- and enough for the label tolerance above, not for the 1e-4 logits tolerance of FP32 mode. */
#define UKBB_MODE_FP16X2  6

#define UKBB_N_CONV 21      /* 13 encoder 3x3 + 5 same_dim 1x1 + fc0 + fc1 + logits */
#define UKBB_MAX_CLASS 8

typedef struct ukbb_fcn ukbb_fcn;

/* One tf.layers.conv2d (+ tf.layers.batch_normalization) of build_FCN, host pointers.
 * kernel: TF HWIO [ksize][ksize][cin][cout] float32.  gamma/beta/mean/variance: [cout]
 * or all NULL (logits layer).  bias: [cout] or NULL. */
typedef struct {
    const float* kernel;
    int ksize, cin, cout, stride;
    const float* gamma;
    const float* beta;
    const float* moving_mean;
    const float* moving_variance;
    const float* bias;
} ukbb_conv_weights;

typedef struct {
    int n_conv;                       /* must be UKBB_N_CONV */
    const ukbb_conv_weights* conv;    /* graph-creation order (network.py:179-229) */
    float bn_eps;                     /* 1e-3 */
} ukbb_fcn_weights;

/* Build an engine on `device`; folds BN into per-channel scale/shift, re-lays the
 * kernels out for the device.  The caller keeps ownership of `w`. */
int ukbb_fcn_create(const ukbb_fcn_weights* w, int n_class, int device, int mode, ukbb_fcn** out);
void ukbb_fcn_destroy(ukbb_fcn* h);

/* image: device float32 [n][y2][x2], already rescaled and zero-padded; x2, y2 multiples of 16.
 * labels: device uint8 [n][y][x] = argmax over the FP32 softmax (first index on ties),
 *         cropped at (x_pre, y_pre); required.
 * logits, prob: optional device float32 [n][y2][x2][n_class] (NULL = not produced; the
 *         reference fetches prob and discards it). */
int ukbb_fcn_forward(ukbb_fcn* h, const float* image, int n, int x2, int y2,
                     int x_pre, int y_pre, int x, int y,
                     uint8_t* labels, float* logits, float* prob, void* stream);

/* vol: device float32, n_slices*y*x voxels in NIfTI order.  Computes the exact numpy
 * 'linear' percentiles q_lo / q_hi (in percent) over ALL voxels, then writes
 * out[n][y2][x2] = float32((double(clip(v)) - vl) / (vh - vl)) inside the image and 0 in
 * the padding.  vl_vh: optional device double[2] receiving (vl, vh).  If clip_in_place
 * is non-zero `vol` itself is clipped like the reference does to its input array. */
int ukbb_fcn_preprocess(ukbb_fcn* h, float* vol, long long n_slices, int x, int y,
                        double q_lo, double q_hi, int x2, int y2, int x_pre, int y_pre,
                        float* out, double* vl_vh, int clip_in_place, void* stream);

/* The second half of ukbb_fcn_preprocess with thresholds the caller already has: out[n][y2][x2] =
 * float32((double(clip(v, vl, vh)) - vl) / (vh - vl)), zero padding.  For a BLOCK of slices of a sequence whose percentiles were
 * taken over the whole sequence elsewhere -- one subject split over several GPUs (the slices are independent once
 * (vl, vh) of common/image_utils.py:70-77 are known; the exchange is 16 bytes, done by the host). */
int ukbb_fcn_rescale(ukbb_fcn* h, float* vol, long long n_slices, int x, int y, double vl, double vh,
                     int x2, int y2, int x_pre, int y_pre, float* out, int clip_in_place, void* stream);

/* Whole-subject call on HOST buffers (pinned for full speed): vol [T][Z][Y][X] float32 in,
 * labels [T][Z][Y][X] uint8 out, vl_vh host double[2] out (may be NULL),
 * counts host int64 [n_slices][n_class] out (may be NULL).  Asynchronous on `stream`:
 * the outputs are valid after the stream (or ukbb_fcn_sync) has been synchronised.
 * `slot` (0 or 1) selects one of two device staging buffers so consecutive subjects overlap. */
int ukbb_fcn_segment_host(ukbb_fcn* h, const float* vol, int x, int y, int z, int t,
                          double q_lo, double q_hi, uint8_t* labels, double* vl_vh,
                          long long* counts, int slot, void* stream);

/* Per-slice class pixel counts of the most recent ukbb_fcn_forward on this handle:
 * counts: device int64 [n][n_class] (cropped region only). */
int ukbb_fcn_class_counts(ukbb_fcn* h, long long* counts, int n, void* stream);

/* Make `stream` wait for every outstanding read-back of ukbb_fcn_segment_host, so that an
 * event recorded on it afterwards brackets the whole H2D -> compute -> D2H chain. */
int ukbb_fcn_join(ukbb_fcn* h, void* stream);

/* Block the host until all work of this handle's device has finished. */
int ukbb_fcn_sync(ukbb_fcn* h);

/* Test hook (tensor-core modes): run conv layer `layer` (graph order: 1..19 in the BF16 / FP16 modes, the 3x3 layers
 * 1..12 in the x3 modes) alone on a device 16-bit NHWC tensor in [n][hi][wi][cin] -> out [n][ho][wo][cout] (rows = Y,
 * columns = X), synchronously.  In the x3 modes both tensors have two planes: [2][n][..] = hi plane | lo plane.
 * level_out selects the tile shape used for resolution level 0..4. */
int ukbb_fcn_debug_conv(ukbb_fcn* h, int layer, const void* in_bf16, int n, int hi, int wi, int level_out,
                        void* out_bf16, void* stream);

/* Test hook (tensor-core modes): copy the first n_elems elements of an intermediate tensor of the most recent
 * ukbb_fcn_forward out as float32 (hi + lo in the x3 modes).  which: 0 / 1 = encoder ping / pong buffer of resolution
 * `level` ([n][h>>level][w>>level][16 << level]; level outputs are pong, ping, ping, ping, ping for levels 0..4),
 * 2 = t_level = fc0 column block applied to same_dim_level ([n][h>>level][w>>level][64], level 1..4). */
int ukbb_fcn_debug_read(ukbb_fcn* h, int which, int level, float* out_f32, long long n_elems, void* stream);

/* Test hook: per-handle switches read by later calls.  bit 0: ukbb_fcn_preprocess / ukbb_fcn_segment_host always take the generic
 * three-pass radix select instead of the integer fast path (the tests compare the two bit for bit). */
int ukbb_fcn_debug_flags(ukbb_fcn* h, int flags);

/* Kernel timer used by bench.py for the roofline line: when enabled, every launch of the fused head kernel
 * (the dominant kernel of the forward) is bracketed by CUDA events on the launching stream.
 * ukbb_fcn_kernel_timer_read synchronises the device, returns the summed duration (ms) and the number of
 * launches recorded since the last read, and clears the list. */
int ukbb_fcn_kernel_timer(ukbb_fcn* h, int enable);
int ukbb_fcn_kernel_timer_read(ukbb_fcn* h, double* total_ms, long long* launches);

/* Introspection used by bench.py: kernels launched by this handle since creation. */
long long ukbb_fcn_launch_count(const ukbb_fcn* h);
int ukbb_fcn_mode(const ukbb_fcn* h);
int ukbb_fcn_n_class(const ukbb_fcn* h);

/* Connected-component statistics of label maps for the quality-control gates (common/cardiac_utils.py:77-166, 1616-1652;
 * common/image_utils.py:227-249 get_largest_cc / remove_small_cc).  labels: device uint8 [n_slices][y][x] (x * y < 65535);
 * classes: HOST int[n_classes] label values; connectivity 1 = faces (scipy.ndimage.label default), 2 = faces + corners
 * (skimage connectivity=2 in the slice plane).  stats: device int32 [n_slices][n_classes][6] =
 * {area, components, components with area > thres, largest area, first pixel index of the largest component (-1 if none),
 *  area kept by remove_small_cc(thres)}.  Asynchronous on `stream`; needs no engine handle. */
int ukbb_cc_stats(const uint8_t* labels, int n_slices, int x, int y, const int* classes, int n_classes, int connectivity,
                  int thres, int* stats, void* stream);

/* ---- Aortic cine segmentation: UNet + bidirectional ConvLSTM (common/network_ao.py:18-64, 255-319;
 * common/deploy_network_ao.py:130-183).  FP32 CUDA-core kernels.  ukbb_ao_create <- import_meta_graph + saver.restore
 * (deploy_network_ao.py:59-60); ukbb_ao_segment <- the whole window loop of deploy_network_ao.py:130-183 (one sess.run per
 * frame there) incl. the weighted overlap-add, the argmax (:186) and the crop. */
typedef struct ukbb_ao ukbb_ao;

typedef struct {
    int n_level;                              /* resolution levels (5), two conv blocks per level (train_network_ao.py:284) */
    const ukbb_conv_weights* down;            /* 2 * n_level encoder conv + BN layers, graph order (network_ao.py:29-41) */
    const ukbb_conv_weights* up_transpose;    /* n_level - 1 transposed convs, levels n_level-2 .. 0: kernel in TF's conv2d_transpose
                                                 layout [3][3][cout][cin], cin = 2 cout, stride 2, + BN (network_ao.py:50-51) */
    const ukbb_conv_weights* up;              /* 2 * (n_level - 1) decoder conv + BN layers, levels n_level-2 .. 0 (:53-54) */
    const float* lstm_kernel[2];              /* forward, backward Conv2DLSTMCell kernel [3][3][f0 + n_hidden][4 n_hidden] (:281-297) */
    const float* lstm_bias[2];                /* [4 n_hidden], gate order (i, j, f, o) */
    const float* out_kernel;                  /* [2 n_hidden][n_class] (network_ao.py:310) */
    const float* out_bias;                    /* [n_class] */
    int n_hidden, n_class;
    float bn_eps;
} ukbb_ao_weights;

int ukbb_ao_create(const ukbb_ao_weights* w, int device, ukbb_ao** out);
void ukbb_ao_destroy(ukbb_ao* h);
/* image: device float32 [n_frames][y2][x2], z-scored (image_utils.py:60-67) and zero-padded (256 x 256 in the reference,
 * deploy_network_ao.py:104-107); labels: device uint8 [n_frames][y][x]; prob: optional device float32 [n_frames][y][x][n_class]
 * = the averaged window probabilities of deploy_network_ao.py:176-180.  weight_R / weight_r: the window flags (5 / 0.1);
 * the time window is 2 weight_R - 1 frames with circular indexing, time_step 1. */
int ukbb_ao_segment(ukbb_ao* h, const float* image, int n_frames, int x2, int y2, int x_pre, int y_pre, int x, int y,
                    int weight_R, double weight_r, uint8_t* labels, float* prob, void* stream);
long long ukbb_ao_launch_count(const ukbb_ao* h);

/* Host helper (no GPU): Castagnoli CRC used by the TF checkpoint bundle reader. */
uint32_t ukbb_crc32c(const void* data, size_t n);

const char* ukbb_last_error(void);
const char* ukbb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* UKBB_FCN_H_ */
