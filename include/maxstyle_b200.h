/*
 * maxstyle_b200.h -- C ABI of the B200-native MaxStyle feature-style layer.
 *
 * This is the drop-in boundary for ONE hot path of cherise215/MaxStyle: the layer in
 * src/advanced/maxstyle.py (forward :140-189, the backward autograd derives from it, and
 * the optimiser step its caller applies at
 * src/models/advanced_triplet_recon_segmentation_model.py:537,562).  The reference has no
 * native code or FFI for this path; these entry points are what a
 * binding for it would call, and maxstyle_b200/_lib.py binds them with ctypes.
 *
 * Conventions
 *  - Plain C: pointers + sizes, no torch / C++ types.  The caller owns every buffer; the
 *    library never allocates or frees device memory and keeps no mutable global state.
 *  - All work is enqueued on the caller's `stream`; no host synchronisation, no default
 *    stream use, graph-capturable.  Re-entrant and thread-safe (autograd calls the backward
 *    from its own thread).  The caller selects the device (cudaSetDevice / device guard).
 *  - Return value: MAXSTYLE_OK or an error code (maxstyle_strerror()).  No exceptions, no abort.
 *  - `workspace` must hold maxstyle_workspace_bytes() bytes, be 256-byte aligned and be
 *    ZERO-FILLED ONCE by the caller before its first use; every call leaves the counters in
 *    it zeroed again, so it can be reused by later calls on the same stream.
 *  - Style tables are fp32 and row-major [rows, C]: mu / sig / scale / shift, gamma_noise /
 *    beta_noise ([N,C,1,1] in the reference) ; lmda is [N] ([N,1,1,1] in the reference);
 *    gamma_std / beta_std are [C] ([1,C,1,1] in the reference); perm is int64 [N_global].
 *  - Global-batch (multi-GPU) extension: a rank holds rows [row_offset, row_offset+N) of a
 *    global batch of N_global samples; `mu_all`/`sig_all` are [N_global, C] and perm indexes
 *    the global batch.  Single GPU: N_global == N, row_offset == 0.
 *  - `table_ld` is the number of floats between consecutive rows of mu_all / sig_all (>= C).
 *    ld == C for two separate [N_global, C] arrays; ld == 2C with sig_all == mu_all + C for one
 *    interleaved [N_global, 2C] buffer (the shape one all-gather of per-rank blocks produces).
 */
#ifndef MAXSTYLE_B200_H_
#define MAXSTYLE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* opaque: a cudaStream_t passed as void* so that the header needs no CUDA include */
typedef void* maxstyle_stream_t;

enum {
    MAXSTYLE_OK = 0,
    MAXSTYLE_ERR_BAD_ARG = 1,      /* null pointer / non-positive size / inconsistent rows   */
    MAXSTYLE_ERR_UNSUPPORTED = 2,  /* dtype / layout / shape this build has no kernel for    */
    MAXSTYLE_ERR_WORKSPACE = 3,    /* workspace missing, too small or misaligned             */
    MAXSTYLE_ERR_CUDA = 4,         /* cudaGetLastError() after a launch was not cudaSuccess  */
    MAXSTYLE_ERR_NO_DEVICE = 5,    /* no sm_100 device / wrong architecture                  */
    MAXSTYLE_ERR_TIMEOUT = 6       /* maxstyle_workspace_status: a device-side wait gave up  */
};

enum { MAXSTYLE_F32 = 0, MAXSTYLE_BF16 = 1 };           /* element type of x / y / dy / dx      */
enum { MAXSTYLE_NCHW = 0, MAXSTYLE_NHWC = 1 };          /* memory layout of x / y / dy / dx     */

/* flags */
enum {
    MAXSTYLE_MIX_STYLE = 1,         /* maxstyle.py:172-176 (else :177-179)                     */
    MAXSTYLE_NO_NOISE = 2,          /* maxstyle.py:181-182 (else :183-185)                     */
    MAXSTYLE_COMPUTE_BATCH_STD = 4, /* first forward: fill gamma_std/beta_std (maxstyle.py:165-168);
                                       otherwise they are inputs (the reference's cached values) */
    MAXSTYLE_NO_CLAMP = 8           /* use lmda as given instead of clamp(lmda, 0, 1): the MixStyle layer
                                       (src/advanced/mixstyle.py:95-96) extrapolates with lmda outside [0,1] */
};

/* sweep flags -- performance only: how a streaming kernel walks its tensor and what it tells L2.
 * The three streaming kernels of one fwd+bwd touch x three times; a caller that runs them back to
 * back alternates the direction so that each kernel starts where the previous one stopped (those
 * lines are still in the 126 MB L2) and marks x "keep" while another kernel will re-read it.
 * MaxStyle.forward/backward use  stats: X_KEEP;  apply: REVERSE | X_KEEP;  bwd: X_STREAM.
 * REVERSE changes the summation order of stats / bwd (last-bit differences); a fixed choice is
 * run-to-run deterministic. */
enum {
    MAXSTYLE_SWEEP_REVERSE = 1,    /* walk each CTA's slice from its high end                         */
    MAXSTYLE_SWEEP_X_KEEP = 2,     /* loads of x: L2 evict-last (a later kernel re-reads x)           */
    MAXSTYLE_SWEEP_X_STREAM = 4,   /* loads of x: L2 evict-first (last use of x)                      */
    MAXSTYLE_SWEEP_IO_NORMAL = 8,  /* y / dy / dx: default L2 policy instead of evict-first            */
    MAXSTYLE_SWEEP_NO_FUSED = 16,  /* maxstyle_fwd (stats_sweep): always take the two-pass path        */
    MAXSTYLE_SWEEP_NO_RESIDENT = 32, /* maxstyle_fwd (stats_sweep): skip the shared-memory-resident kernel */
    MAXSTYLE_SWEEP_FORCE_WINDOW = 64, /* maxstyle_fwd (stats_sweep): take the L2-window kernel whenever the
                                        shape qualifies, even where the two-pass path is faster (tests)   */
    MAXSTYLE_SWEEP_FORCE_RESIDENT = 128, /* ... likewise for the shared-memory-resident kernel              */
    MAXSTYLE_SWEEP_NO_RING = 256,   /* maxstyle_fwd (stats_sweep): skip the TMA-ring version of the L2-window kernel */
    MAXSTYLE_SWEEP_FORCE_RING = 512, /* ... take it whenever the shape qualifies                              */
    MAXSTYLE_SWEEP_NO_CLUSTER = 1024, /* maxstyle_fwd / maxstyle_fwd_p2p (stats_sweep): skip the cluster-resident kernel */
    MAXSTYLE_SWEEP_FORCE_CLUSTER = 2048, /* ... take it whenever the shape qualifies                           */
    MAXSTYLE_SWEEP_CLUSTER_SIZE_SHIFT = 12,   /* bits 12-15: CTAs per cluster for the cluster-resident kernel (1, 2, 4, 8; 0 = chosen by the library) */
    MAXSTYLE_SWEEP_CLUSTER_STAGES_SHIFT = 16, /* bits 16-18: cap on its stages per CTA (0 = as many as fit)     */
    MAXSTYLE_SWEEP_CLUSTER_PIECES_SHIFT = 19, /* bits 19-24: pieces a plane is cut into, cluster-resident and paired kernels (0 = chosen by the library) */
    MAXSTYLE_SWEEP_NO_PAIR = 1 << 25,         /* maxstyle_fwd / maxstyle_fwd_p2p (stats_sweep): skip the paired (piece-owning) kernel */
    MAXSTYLE_SWEEP_FORCE_PAIR = 1 << 26       /* ... take it whenever the shape qualifies                           */
};

/* activation fused in front of the layer: the kernels take the PRE-activation tensor z and apply the
 * activation to every value as it is loaded, x = act(z); the backward returns dz */
enum {
    MAXSTYLE_PRE_NONE = 0,
    MAXSTYLE_PRE_LEAKY_RELU = 1,  /* x = z > 0 ? z : slope*z -- the LeakyReLU(0.2) that ends res_up_family
                                     (src/models/ebm/encoder_decoder.py:337-357) in front of layers 0-4            */
    MAXSTYLE_PRE_SIGMOID = 2      /* x = 1/(1+exp(-z)) -- the decoder's `last_act` in front of layer 5 (:624-630)   */
};

/* optimiser step fused into the backward epilogue (north_star item 4) */
enum {
    MAXSTYLE_STEP_NONE = 0,
    MAXSTYLE_STEP_ADAM = 1,  /* torch.optim.Adam semantics, the reference's choice (model:537)  */
    MAXSTYLE_STEP_SIGN = 2   /* p -= lr*sign(g) (or += with maximize): sign-gradient step       */
};

typedef struct maxstyle_step {
    int32_t mode;            /* MAXSTYLE_STEP_*                                                   */
    int32_t maximize;        /* 0: descend the gradient handed to the backward (the reference
                                backpropagates -CE, model:555);  1: ascend it                    */
    double lr, beta1, beta2, eps; /* doubles: bias corrections are computed like torch does, in fp64 */
    int32_t t;               /* 1-based step number for Adam's bias correction; ignored when
                                step_dev != NULL                                                  */
    int32_t update_noise;    /* gamma_noise / beta_noise are learnable (noise_learnable)          */
    int32_t update_mix;      /* lmda is learnable (mix_learnable)                                 */
    int32_t reserved;
    int32_t* step_dev;       /* optional device counter: the kernel uses *step_dev+1 as t and
                                increments it once per call (CUDA-graph replay friendly)         */
    float* gamma_noise;      /* [N,C] updated in place                                            */
    float* beta_noise;       /* [N,C]                                                             */
    float* lmda;             /* [N]                                                               */
    float* gamma_m; float* gamma_v;   /* Adam exp_avg / exp_avg_sq, same shapes as the parameter */
    float* beta_m;  float* beta_v;
    float* lmda_m;  float* lmda_v;
} maxstyle_step_t;

/* Library / build identification, e.g. "maxstyle_b200 0.1 sm_100a". */
const char* maxstyle_version(void);
const char* maxstyle_strerror(int code);

/* Bytes of scratch the calls below need for this problem (counters + per-tile partials). */
size_t maxstyle_workspace_bytes(int N, int C, int H, int W, int dtype, int layout);

/* Kernel 1 -- instance statistics (replaces x.mean / x.var / sqrt, maxstyle.py:157-159).
 * One pass over x; per-(n,c) plane mean and sqrt(unbiased var + eps) by Welford/Chan
 * merging.  Writes rows [row_offset, row_offset+N) of mu_all / sig_all. */
int maxstyle_stats(const void* x, float* mu_all, float* sig_all, int table_ld, int row_offset,
                   int N, int C, int H, int W, int dtype, int layout, float eps, int sweep,
                   void* workspace, size_t workspace_bytes, maxstyle_stream_t stream);

/* Table step (replaces maxstyle.py:165-185 on the [N,C] tables): optional batch std,
 * clamp(lmda), partner gather through perm, lerp, noise perturbation.  Produces for the
 * local rows  scale = A/sig  and  shift = B  with  A = sig_mix + gamma_noise*gamma_std,
 * B = mu_mix + beta_noise*beta_std, so that  y = (x - mu)*scale + shift. */
int maxstyle_tables(const float* mu_all, const float* sig_all, int table_ld, int N_global, int row_offset,
                    int N, int C,
                    const int64_t* perm, const float* lmda, const float* gamma_noise, const float* beta_noise,
                    float* gamma_std, float* beta_std, int flags,
                    float* scale, float* shift, maxstyle_stream_t stream);

/* Multi-GPU: the exchange of the (mu | sig) rows AND the table step in ONE kernel over NVLink peer memory -- the fused
 * form of "all-gather, then maxstyle_tables" (no reference counterpart: the reference is single-GPU; north_star's one
 * collective on the path).  `peers` is a DEVICE array of `world` addresses: entry r is rank r's exchange buffer of
 * maxstyle_p2p_bytes(N, C, world) bytes, mapped into this process (symmetric memory / CUDA IPC; the caller owns and maps it,
 * zero-filled once).  Rank `rank` must have filled rows [rank*N, (rank+1)*N) of mu_all / sig_all (maxstyle_stats); on return
 * (stream-ordered) all N_global = N*world rows are present, exactly as after an all-gather, and scale / shift (and on the
 * first forward gamma_std / beta_std) are computed from them.  CTA c pushes channel c of this rank's rows into every
 * peer's buffer as 8-byte {value, epoch} words with single stores over NVLink (the epoch tag is the arrival flag: no
 * fence, no round trip), spins on the words of channel c in its own buffer and copies them into the table; `epoch` (device counter,
 * starts at 0, advanced by the kernel) numbers the exchanges so the call replays from a CUDA graph with fixed arguments;
 * `done` is a zeroed device word; if a peer does not publish within ~20 s `error` is set to 1 and the kernel traps (the launch
 * fails).  Every rank must make the same sequence of calls. */
size_t maxstyle_p2p_bytes(int N, int C, int world);
int maxstyle_tables_p2p(const uint64_t* peers, int rank, int world, uint32_t* epoch, uint32_t* done, int* error,
                        float* mu_all, float* sig_all, int table_ld, int N_global, int row_offset, int N, int C,
                        const int64_t* perm, const float* lmda, const float* gamma_noise, const float* beta_noise,
                        float* gamma_std, float* beta_std, int flags, float* scale, float* shift, maxstyle_stream_t stream);

/* Rank barrier over the exchange buffers (one warp per rank, tagged words pushed to every peer): enqueue it right before
 * maxstyle_fwd_p2p so that all ranks start that kernel together -- the launch skew between ranks is then waited out before
 * any streaming instead of inside the forward's L2 window.  `bar_epoch` is a device counter of its own (starts at 0). */
int maxstyle_rank_barrier(const uint64_t* peers, int rank, int world, int N, int C, uint32_t* bar_epoch, int* error,
                          maxstyle_stream_t stream);

/* Multi-GPU whole forward in ONE kernel: the paired forward of maxstyle_fwd (every CTA owns a piece of a plane for both of its
 * passes) whose piece-0 owners push their plane's (mu, sig) as 8-byte {value, tag} words into every peer's inbox -- the same
 * peer-memory inboxes and epoch counter as maxstyle_tables_p2p (the two calls may be mixed on one set of buffers as long as
 * every rank makes the same sequence of calls) -- and whose CTAs poll only their own inbox for the partner rows.  x is read from
 * HBM once, y written once.  When a channel's pieces do not fit the grid the samples are taken in cycle order of the global
 * permutation (N <= 1024 per rank, N_global <= 2048); the first forward of a layer (MAXSTYLE_COMPUTE_BATCH_STD) needs the whole
 * channel in the grid.  With 2 ranks the L2-window forward is the second choice.  Returns MAXSTYLE_ERR_UNSUPPORTED without
 * launching anything when the shape does not qualify (the caller then runs maxstyle_stats -> maxstyle_tables_p2p ->
 * maxstyle_apply); every rank sees the same answer for the same shape and flags.  A peer that does not publish within ~20 s
 * raises the workspace error flag and traps: the launch fails (maxstyle_workspace_status reads the flag). */
int maxstyle_fwd_p2p(const void* x, void* y, float* mu_all, float* sig_all, int table_ld, int N_global, int row_offset,
                     const int64_t* perm, const float* lmda, const float* gamma_noise, const float* beta_noise,
                     float* gamma_std, float* beta_std, float* scale, float* shift,
                     int N, int C, int H, int W, int dtype, int layout, int flags, float eps, int stats_sweep,
                     const uint64_t* peers, int rank, int world, uint32_t* epoch,
                     void* workspace, size_t workspace_bytes, maxstyle_stream_t stream);

/* Kernel 2 -- apply (replaces the normalise + affine chain, maxstyle.py:161,181-185). */
int maxstyle_apply(const void* x, void* y, const float* mu_all, int table_ld, int row_offset, const float* scale,
                   const float* shift, int N, int C, int H, int W, int dtype, int layout, int sweep,
                   maxstyle_stream_t stream);

/* Whole forward on one GPU (N_global == N), one call.  Replaces MaxStyle.forward's active path
 * (maxstyle.py:157-185).  Three implementations, same results up to summation order, tried in this order:
 *  - resident (16-byte-multiple planes of 16 KB .. ~108 KB so that two CTAs share an SM, tensors >= 64 MB, 2 <= N <=
 *    number of co-resident CTAs; planes up to ~220 KB with MAXSTYLE_SWEEP_FORCE_RESIDENT): ONE persistent
 *    kernel; each plane is brought into shared memory once by TMA bulk copies, its moments are taken there, the
 *    CTA exchanges (mu, sig) with its mixing partner's CTA through global flags, and y is written from shared
 *    memory while the next plane is already loading -- HBM sees x once and y once, L2 is only passed through;
 *  - L2 window, streamed (same rule as the next one): the ordered statistics/apply queue described below, but x moves
 *    through a ring of TMA bulk copies that one producer thread keeps full across item boundaries, and the
 *    consumer warps work out of shared memory;
 *  - L2 window (when the shape qualifies and it pays: planes >= 64 KB, a channel <= 8 MB: planes are 16-byte multiples of at least 8-16 KB, 2 <= N <= 1024,
 *    one channel of x is at most 16 MB, so the L2 window holds >= 2 channels): ONE persistent kernel working through an ordered queue of
 *    statistics and apply items, channel-major, with the apply items a ~32 MB window behind the
 *    statistics items, so the second read of x comes out of L2 -- HBM sees x once and y once;
 *  - two-pass: maxstyle_stats -> maxstyle_tables -> maxstyle_apply (x read from HBM twice). */
int maxstyle_fwd(const void* x, void* y, float* mu, float* sig,
                 const int64_t* perm, const float* lmda, const float* gamma_noise, const float* beta_noise,
                 float* gamma_std, float* beta_std, float* scale, float* shift,
                 int N, int C, int H, int W, int dtype, int layout, int flags, float eps,
                 int stats_sweep, int apply_sweep,
                 void* workspace, size_t workspace_bytes, maxstyle_stream_t stream);

/* maxstyle_fwd with its neighbours fused in (NCHW only):
 *  - `pre_op` / `pre_param`: `z` is the tensor BEFORE the activation that precedes the layer in the reference's decoders
 *    (MAXSTYLE_PRE_LEAKY_RELU with its negative slope, or MAXSTYLE_PRE_SIGMOID); the layer sees x = act(z) without x ever
 *    being written to memory -- one full-tensor pass (read z, write x) less than activation + layer;
 *  - `y_min` / `y_max` (optional, both or neither; [N*C]): per-plane minimum and maximum of y, collected while y is written,
 *    as order-preserving unsigned integers reduced with atomicMin / atomicMax -- the caller initialises them to 0xffffffff
 *    and 0; maxstyle_rescale turns them into the reference's rescale_intensity (src/common_utils/basic_operations.py:257-281,
 *    applied to the loop's output at advanced_triplet_recon_segmentation_model.py:868-869) in one more pass.
 * Runs as the paired kernel, else as statistics -> tables -> apply.  maxstyle_bwd_act is the matching backward: it recomputes
 * x = act(z) on the fly and writes dz = dy * A/sig * act'(z). */
int maxstyle_fwd_act(const void* z, void* y, float* mu, float* sig,
                     const int64_t* perm, const float* lmda, const float* gamma_noise, const float* beta_noise,
                     float* gamma_std, float* beta_std, float* scale, float* shift,
                     int N, int C, int H, int W, int dtype, int layout, int flags, float eps,
                     int stats_sweep, int apply_sweep, int pre_op, float pre_param, uint32_t* y_min, uint32_t* y_max,
                     void* workspace, size_t workspace_bytes, maxstyle_stream_t stream);
int maxstyle_bwd_act(const void* dy, const void* z, void* dz,
                     const float* mu_all, const float* sig_all, int table_ld, int N_global, int row_offset,
                     const float* scale, const int64_t* perm, const float* lmda,
                     const float* gamma_std, const float* beta_std, int flags,
                     float* d_gamma, float* d_beta, float* d_lmda,
                     const maxstyle_step_t* step,
                     int N, int C, int H, int W, int dtype, int layout, int sweep, int pre_op, float pre_param,
                     void* workspace, size_t workspace_bytes, maxstyle_stream_t stream);
/* out = (y - min) / (max - min + eps) * (new_max - new_min) + new_min per (n,c) plane, min / max as collected by maxstyle_fwd_act. */
int maxstyle_rescale(const void* y, const uint32_t* y_min, const uint32_t* y_max, void* out, float new_min, float new_max, float eps,
                     int N, int C, int H, int W, int dtype, maxstyle_stream_t stream);

/* Number of kernels maxstyle_fwd launches for this shape on the current device: 1 (resident or L2 window), 3 (two-pass:
 * stats, tables, apply), 0 for an unsupported shape.  Assumes 16-byte aligned x and y. */
int maxstyle_fwd_kernels(int N, int C, int H, int W, int dtype, int layout, int stats_sweep);

/* How the single-kernel forwards would run this shape on the current device in their steady state (diagnostics; no launch).
 * With MAXSTYLE_SWEEP_FORCE_CLUSTER in stats_sweep, the cluster-resident kernel: out[0..9] = CTAs per cluster, stages per CTA,
 * co-resident clusters, bytes of a piece per CTA, chunk bytes, chunks per part, dynamic shared memory per CTA, SMs, pieces per
 * plane, 1 if the samples are visited in cycle order of perm.  Otherwise the paired kernel: out[2] = CTAs, out[3] = bytes per
 * piece, out[7] = SMs, out[8] = pieces per plane, out[9] = cycle order, out[10] = 1.  `out` holds 12 ints.
 * MAXSTYLE_ERR_UNSUPPORTED when the kernel cannot take the shape. */
int maxstyle_fwd_geometry(int N, int C, int H, int W, int dtype, int stats_sweep, int* out);

/* Debug / test helper, the only entry point that synchronises: waits for `stream` and returns
 * MAXSTYLE_ERR_TIMEOUT if a device-side wait of the fused forward gave up since the workspace
 * was zero-filled (cannot happen unless the co-residency guarantee of the cooperative launch is broken). */
int maxstyle_workspace_status(const void* workspace, size_t workspace_bytes, int N, int C, int H, int W, int dtype, int layout,
                              maxstyle_stream_t stream);

/* Kernel 3 -- backward (replaces the autograd graph of maxstyle.py:157-185).
 * One sweep over dy and x:  dx = dy*scale (skipped when dx == NULL),  dA = sum dy*(x-mu)/sig,
 * dB = sum dy;  epilogue per sample:  d_gamma = dA*gamma_std,  d_beta = dB*beta_std,
 * d_lmda = [0<=lmda<=1] * sum_c (dA*(sig[perm]-sig) + dB*(mu[perm]-mu)),  then the optional
 * fused optimiser step.  d_gamma / d_beta / d_lmda may each be NULL (gradient not wanted). */
int maxstyle_bwd(const void* dy, const void* x, void* dx,
                 const float* mu_all, const float* sig_all, int table_ld, int N_global, int row_offset,
                 const float* scale, const int64_t* perm, const float* lmda,
                 const float* gamma_std, const float* beta_std, int flags,
                 float* d_gamma, float* d_beta, float* d_lmda,
                 const maxstyle_step_t* step,
                 int N, int C, int H, int W, int dtype, int layout, int sweep,
                 void* workspace, size_t workspace_bytes, maxstyle_stream_t stream);

/* Pixel-wise cross entropy on NCHW logits with an int64 label map [N,H,W] -- the loss whose gradient enters the inner
 * style-optimisation loop (replaces cross_entropy_2D, src/models/custom_loss.py:1043-1105, label-map branch :1069-1078):
 *   loss = -(1/D) sum_p mask[p] * weight[t_p] * log_softmax(logits[:, :, p])[t_p],  D = N*H*W if size_average else 1.
 * `weight` ([C], already normalised by the caller as the reference does: w / sum(w) * C) and `mask` ([N*H*W]) may be NULL.
 * Labels equal to -100 are ignored (F.nll_loss's default ignore_index, inherited by the reference); any other label outside
 * [0, C) stops the kernel with a device-side trap (the reference's F.nll_loss raises a device assert).  One kernel each way;
 * the forward's sum is deterministic.  `workspace`: maxstyle_ce2d_workspace_bytes() bytes, 256-byte aligned, zero-filled once.
 * The backward takes the upstream gradient of the scalar loss from device memory (`dloss`, 1 float). */
size_t maxstyle_ce2d_workspace_bytes(int N, int C, int H, int W);
int maxstyle_ce2d_fwd(const void* logits, const int64_t* target, const float* weight, const float* mask, float* loss,
                      int N, int C, int H, int W, int dtype, int size_average, void* workspace, size_t workspace_bytes,
                      maxstyle_stream_t stream);
int maxstyle_ce2d_bwd(const void* logits, const int64_t* target, const float* weight, const float* mask, const float* dloss,
                      void* dlogits, int N, int C, int H, int W, int dtype, int size_average, maxstyle_stream_t stream);

/* The same loss AND its gradient in ONE kernel (the inner loop always back-propagates it: model:555-561), both branches of
 * cross_entropy_2D: a label map (`labels`, int64 [N,H,W]; custom_loss.py:1069-1078) or a soft target (`soft_target`,
 * [N,C,H,W] of the logits' dtype; logits unless `soft_is_probability`; custom_loss.py:1079-1102) -- exactly one of the two.
 * `dlogits` (optional) receives d loss / d logits and `dsoft_target` (optional, soft branch) d loss / d target, both for an upstream
 * gradient of 1; maxstyle_ce2d_scale multiplies a gradient by the actual upstream gradient, read from device memory.
 * C <= 8 (more: MAXSTYLE_ERR_UNSUPPORTED, use maxstyle_ce2d_fwd / _bwd).  A label outside [0, C) other than -100 stops the
 * kernel with a device-side trap, as F.nll_loss's device assert does in the reference. */
int maxstyle_ce2d_fwd_grad(const void* logits, const int64_t* labels, const void* soft_target, int soft_is_probability,
                           const float* weight, const float* mask, float* loss, void* dlogits, void* dsoft_target,
                           int N, int C, int H, int W, int dtype, int size_average, void* workspace, size_t workspace_bytes,
                           maxstyle_stream_t stream);
int maxstyle_ce2d_scale(void* grad, const float* scale, int64_t count, int dtype, maxstyle_stream_t stream);

/* Stand-alone optimiser step on the three parameter tensors (same arithmetic as the fused
 * epilogue; used when the gradients arrive through autograd's .grad instead). */
int maxstyle_step(const float* d_gamma, const float* d_beta, const float* d_lmda,
                  const maxstyle_step_t* step, int N, int C, maxstyle_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MAXSTYLE_B200_H_ */
