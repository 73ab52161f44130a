/*
 * pil.h -- C ABI of the B200-native physics-prior loss ("pil") library.
 *
 * This is the drop-in boundary for the hot path of seemapoudel58/Physics_informed_image_segmentation:
 * the Stage II loss  L = wd*Dice + wb*BCE + lambda_rd*mean((D lap(u) + u(1-u)(u-a))^2)
 *                        + lambda_pf*mean((eps/2)|grad u|^2 + u^2(1-u)^2/eps)
 * evaluated forward and backward on single-channel probability maps.
 *
 * The reference has no FFI of its own (it is pure Python/PyTorch); the entry points below are what a
 * binding for this path binds.  Each one cites the reference interface it replaces
 * (file:line relative to the reference checkout).  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *  - All tensor pointers are DEVICE pointers to contiguous (B,1,H,W) == (B,H,W) row-major maps,
 *    except in the pil_session_* calls, which take HOST pointers.
 *  - `stream` is a cudaStream_t passed as void*.  Every call is stream-ordered, allocates nothing,
 *    never synchronises the host (except pil_session_run, which returns the loss to the host) and is
 *    re-entrant across streams as long as each stream uses its own workspace.
 *  - Return value: PIL_OK (0), a negative PilStatus, or a positive cudaError_t.
 *  - There is no CPU fallback: without a CUDA device every compute call fails with a cudaError_t.
 */
#ifndef PIL_H_
#define PIL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIL_VERSION 100 /* 0.1.0 */

/* Number of doubles in a sums vector (see pil_forward). */
#define PIL_NSUMS 8
/* Number of floats in a loss-report vector (see pil_finalize). */
#define PIL_NOUT 8
/* Number of doubles in a moments vector (see pil_forward_moments). */
#define PIL_NMOMENTS 16

typedef enum PilStatus {
    PIL_OK = 0,
    PIL_ERR_NULL = -1,        /* a required pointer is NULL */
    PIL_ERR_SHAPE = -2,       /* B<1, H<2 or W<2 (reflect padding needs >= 2, src/pde.py:67) */
    PIL_ERR_DTYPE = -3,       /* unsupported dtype combination */
    PIL_ERR_KIND = -4,        /* unknown input kind */
    PIL_ERR_WORKSPACE = -5,   /* workspace too small or misaligned */
    PIL_ERR_DIFFUSION = -6,   /* diffusion_coeff <= 0        (ValueError at src/pde.py:14-15) */
    PIL_ERR_THRESHOLD = -7,   /* reaction_threshold not in (0,1) (ValueError at src/pde.py:16-17) */
    PIL_ERR_EPSILON = -8,     /* epsilon <= 0 while phase_field_weight > 0 (src/pde.py:199-200) */
    PIL_ERR_ALIGNMENT = -9,   /* a pointer is not aligned to its element size */
    PIL_ERR_SESSION = -10,    /* session misuse (batch larger than created for, ...) */
    PIL_ERR_EXCHANGE = -11    /* bad PilExchange descriptor */
} PilStatus;

typedef enum PilDtype {
    PIL_F32 = 0,  /* predictions / logits / targets / gradient */
    PIL_BF16 = 1, /* predictions / logits / targets / gradient (math stays fp32) */
    PIL_U8 = 2    /* targets only: {0,1} masks */
} PilDtype;

/* What `x` holds.  The reference applies the activation inside UNet.forward (src/unet.py:208-214)
 * and hands probabilities to the loss; fusing it is the "logits" entry. */
typedef enum PilInputKind {
    PIL_X_PROB = 0,           /* x = u in [0,1]; backward writes dL/du  (strict drop-in, src/loss.py:114) */
    PIL_X_LOGITS_SIGMOID = 1, /* x = z, u = sigmoid(z); backward writes dL/dz (src/unet.py:210) */
    PIL_X_LOGITS_TANH = 2     /* x = z, u = (tanh(z)+1)/2; backward writes dL/dz (src/unet.py:211-214) */
} PilInputKind;

/* The knobs of DiceBCEPDELoss.__init__ (src/loss.py:86-96), same meaning, as Python floats (double).
 * CLI: --pde-weight, --phase-field-weight, --diffusion-coeff, --reaction-threshold, --epsilon
 * (main.py:14-43).  A weight of exactly 0 (or less) switches its term off entirely, like the Python
 * `> 0` gates at src/loss.py:150 and :155. */
typedef struct PilParams {
    double dice_weight;
    double bce_weight;
    double pde_weight;
    double phase_field_weight;
    double diffusion_coeff;
    double reaction_threshold;
    double epsilon;
    double smooth;
} PilParams;

struct PilExchange; /* peer-memory exchange descriptor, defined below */

int pil_version(void);
const char* pil_status_string(int status);

/* Constructor-time and call-time validation of the reference (src/pde.py:14-17, :199-200). */
int pil_validate_params(const PilParams* p);

/* Scratch the kernels need for their deterministic two-level reductions.  The caller owns it, must zero ALL of
 * it once (pil_workspace_init, or cudaMemset over pil_workspace_bytes) and may reuse it call after call on the same
 * stream; the kernels leave it ready for the next launch.  (The per-block partial sums are validated by a per-launch
 * tag instead of a memory fence, so stale bytes must not look like a tag: hence the zero.) */
size_t pil_workspace_bytes(int64_t B, int64_t H, int64_t W);
int pil_workspace_init(void* workspace, size_t workspace_bytes, void* stream);

/*
 * Forward: ONE fused kernel.  Replaces DiceBCEPDELoss.forward / DiceBCELoss.forward
 * (src/loss.py:114-162, :36-68) including PDERegularization.compute_loss (src/pde.py:124-145) and
 * compute_phase_field_loss (src/pde.py:180-212), and -- for the logits kinds -- the model's output
 * activation (src/unet.py:208-214).
 *
 * sums (device, PIL_NSUMS doubles, overwritten) for THIS shard of images:
 *   [0] I = sum u*t   [1] P = sum u   [2] T = sum t                    (src/loss.py:134-137)
 *   [3] sum of BCE terms with the log clamp at -100                       (nn.BCELoss, src/loss.py:141)
 *   [4] sum r^2, r = D*lap(u) + u(1-u)(u-a)                               (src/pde.py:117-120,:143)
 *   [5] sum (eps/2)(gx^2+gy^2) + u^2(1-u)^2/eps                           (src/pde.py:203-210)
 *   [6] number of u outside [0,1] or NaN (nn.BCELoss raises on those)
 *   [7] number of pixels B*H*W
 * Under data parallelism the caller all-reduces (SUM) `sums` across ranks and passes the result to
 * pil_finalize / pil_backward.  If loss_out != NULL the kernel also finalises on the spot as if this
 * shard were the whole batch (single-GPU fast path; same layout as pil_finalize).
 */
int pil_forward(const void* x, const void* t, int64_t B, int64_t H, int64_t W,
                int x_dtype, int t_dtype, int x_kind, const PilParams* p,
                double* sums, float* loss_out, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Assemble the scalar loss from (all-reduced) sums on the device: src/loss.py:134-160.
 * n_global <= 0 means "use sums[7]".
 * loss_out (device, PIL_NOUT floats): [0] total  [1] dice_loss  [2] bce  [3] L_rd  [4] L_pf
 * [5] n_invalid  [6],[7] reserved.  If n_invalid > 0 (probabilities outside [0,1] or NaN, on which the
 * reference's nn.BCELoss raises) the total is NaN: stream-ordered code cannot raise, it fails loudly instead.  [1..4] are the four quantities train_epoch re-computes for
 * logging at src/train.py:120-150, served here without a second pass over the maps.
 */
int pil_finalize(const double* sums, int64_t n_global, const PilParams* p, float* loss_out, void* stream);

/*
 * Backward: ONE fused kernel writing grad = upstream * grad_scale * dL/dx (same dtype as x).
 * Replaces autograd through src/loss.py:114-162 / src/pde.py (ConvolutionBackward x3,
 * ReflectionPad2dBackward x2, BinaryCrossEntropyBackward with its 1e-12 clamp, ...)
 * and, for the logits kinds, SigmoidBackward/TanhBackward of src/unet.py:208-214.
 * global_sums: device, the all-reduced sums of pil_forward; n_global: pixels in the GLOBAL batch, or
 * <= 0 to use the all-reduced count global_sums[7] on the device (unequal shards, no host sync).
 * upstream: device pointer to the scalar float gradient of the loss (NULL = 1.0), read on the device
 * so that loss.backward() needs no host sync.  grad_scale: host scalar (world_size correction under
 * DDP-style gradient averaging).
 */
int pil_backward(const void* x, const void* t, void* grad, int64_t B, int64_t H, int64_t W,
                 int x_dtype, int t_dtype, int x_kind, const PilParams* p,
                 const double* global_sums, int64_t n_global, const float* upstream, float grad_scale,
                 void* stream);

/*
 * Training-step split (what `loss = criterion(x, t); loss.backward()` of src/train.py:117,:163 costs
 * when both halves are known to run): the stencils are evaluated ONCE, in the backward kernel.
 *
 *   pil_forward_pointwise     flat streaming kernel, only the sums that need no neighbours:
 *                             sums = {I, P, T, sum bce, 0, sum u^2(1-u)^2/eps, n_invalid, n_pixels}
 *   (data parallel: all-reduce sums here -- the gradient needs the global I, P, T)
 *   pil_backward_accumulate   pil_backward + the two stencil sums of this shard:
 *                             stencil_sums = {0,0,0,0, sum r^2, (eps/2) sum(gx^2+gy^2), 0, 0}
 *                             sums + stencil_sums is exactly what pil_forward returns.  If loss_out is
 *                             given, the last block finalises the loss from global_sums + stencil_sums
 *                             (single shard); data parallel callers all-reduce stencil_sums and call
 *                             pil_finalize instead.
 *   pil_loss_fwd_bwd          the two above back to back for a single shard (upstream gradient 1):
 *                             writes grad, the complete sums and the loss report.
 *   pil_scale_gradient        grad *= *upstream on the device; returns immediately (no memory traffic)
 *                             when *upstream == 1, the loss.backward() case.
 */
int pil_forward_pointwise(const void* x, const void* t, int64_t B, int64_t H, int64_t W,
                          int x_dtype, int t_dtype, int x_kind, const PilParams* p,
                          double* sums, void* workspace, size_t workspace_bytes, void* stream);
int pil_backward_accumulate(const void* x, const void* t, void* grad, int64_t B, int64_t H, int64_t W,
                            int x_dtype, int t_dtype, int x_kind, const PilParams* p,
                            const double* global_sums, int64_t n_global, const float* upstream, float grad_scale,
                            double* stencil_sums, float* loss_out,
                            void* workspace, size_t workspace_bytes, void* stream);
int pil_loss_fwd_bwd(const void* x, const void* t, void* grad, int64_t B, int64_t H, int64_t W,
                     int x_dtype, int t_dtype, int x_kind, const PilParams* p,
                     double* sums, float* loss_out, void* workspace, size_t workspace_bytes, void* stream);
int pil_scale_gradient(void* grad, int dtype, int64_t n, const float* upstream, void* stream);
/* pil_backward that returns at once when *upstream == 1 (decided on the device): `grad` already holds the gradient
 * for a unit upstream (written by pil_loss_fwd_bwd / pil_backward_accumulate) and is recomputed in place, from the
 * maps and in fp32, only for another upstream value.  For bf16 gradients, where rescaling the stored values would
 * round twice. */
int pil_backward_if_scaled(const void* x, const void* t, void* grad, int64_t B, int64_t H, int64_t W,
                           int x_dtype, int t_dtype, int x_kind, const PilParams* p,
                           const double* global_sums, int64_t n_global, const float* upstream, float grad_scale,
                           void* stream);

/*
 * Per-step accuracy metrics folded into the step.  train_epoch / validate compute, between
 * forward and backward, the per-image Dice and IoU of the prediction thresholded at 0.5
 * (src/train.py:153-160 -> src/metrics.py:38-73 compute_dice_score_batch, src/evaluate.py:62-97
 * compute_iou_batch: B Python iterations of ~6 kernels each).  pil_forward_pointwise_metrics IS the
 * pointwise forward (same sums, same exchange push when ex != NULL) and, on the same read of x and t, also
 * leaves three counts per image:
 *   image_counts (device, B x 4 doubles, zeroed by the call): [b][0] sum [u > threshold]*t
 *   [b][1] sum [u > threshold]   [b][2] sum t   [b][3] 0
 * pil_image_metrics turns them into dice[b] = (2I+s)/(P+T+s) and iou[b] = (I+s)/(P+T-I+s) on the device
 * (either output may be NULL).  The boundary-F1 of src/evaluate.py:125-229 (OpenCV contours + distance
 * transform on the host) is not part of this path.
 */
int pil_forward_pointwise_metrics(const void* x, const void* t, int64_t B, int64_t H, int64_t W,
                                  int x_dtype, int t_dtype, int x_kind, const PilParams* p,
                                  double* sums, double* image_counts, float threshold,
                                  void* workspace, size_t workspace_bytes,
                                  const struct PilExchange* ex /* may be NULL */, void* stream);
int pil_image_metrics(const double* image_counts, int64_t B, double smooth,
                      float* dice_out, float* iou_out, void* stream);

/*
 * Boundary-F1 of the thresholded prediction against the mask, per image, on the device: what
 * compute_boundary_f1_batch (src/evaluate.py:196-229 -> compute_boundary_f1 :125-193 -> extract_boundaries :102-122)
 * computes on the host with OpenCV for every image of every training / validation step (src/train.py:156, :259).
 *   boundary(M)  = cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_NONE) drawn with thickness 1: the foreground pixels
 *                  4-adjacent to background that is 4-connected to the image frame (holes contribute nothing);
 *   within tol   = cv2.distanceTransform(DIST_L2, mask 5) <= tol: a boundary pixel of the other map at a 5x5-chamfer
 *                  distance <= tol (tol = 2: 13 offsets); tol = 0 is the reference's exact-match branch.
 * pil_boundary_counts leaves four integers per image (device, B x 4 int64, zeroed by the call):
 *   [b][0] |Bp|  [b][1] |Bt|  [b][2] #{p in Bp within tol of Bt} (tol 0: |Bp & Bt|)  [b][3] #{q in Bt within tol of Bp}
 * pil_boundary_f1 turns them into the reference's float32 F1 per image.  x: probabilities or logits (x_kind), the
 * prediction mask is activation(x) > threshold; the target mask is (uint8)(t * 255) != 0 as in extract_boundaries.
 * 0 <= tolerance <= 6 (the chamfer weights make larger tolerances ambiguous at equality).  Shapes: H, W >= 1.
 * pil_boundary_tolerance_offsets lists the offsets of a tolerance (host; NULL outputs: just the count).
 */
size_t pil_boundary_workspace_bytes(int64_t B, int64_t H, int64_t W);
int pil_boundary_tolerance_offsets(int tolerance, int8_t* dy_out, int8_t* dx_out, int capacity);
int pil_boundary_counts(const void* x, const void* t, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype, int x_kind,
                        float threshold, int tolerance, long long* counts, void* workspace, size_t workspace_bytes,
                        void* stream);
int pil_boundary_f1(const long long* counts, int64_t B, int tolerance, double smooth, float* f1_out, void* stream);

/*
 * The model tail fused with the loss: the U-Net's 1x1 output convolution from its C-channel
 * full-resolution features to one channel (src/unet.py:157, :205), its output activation (:208-214) and the loss.
 *   pil_tail_forward   z = sum_c weight[c] * feat[b,c,h,w] + bias, written once as fp32 logits (B,1,H,W), AND the
 *                      pointwise sums of pil_forward_pointwise on the same registers (sums; pushed to the peers
 *                      when ex != NULL) -- one pass over the features replaces the convolution, the activation
 *                      kernel and the pointwise forward.  x_kind: PIL_X_LOGITS_SIGMOID or PIL_X_LOGITS_TANH.
 *   (then pil_backward_accumulate[_xchg] on logits_out, which yields grad_logits = dL/dlogits and the loss report)
 *   pil_tail_backward  ONE pass over the features: grad_feat[b,c,h,w] = weight[c] * grad_logits[b,h,w] (same dtype
 *                      as feat), grad_weight[c] = sum grad_logits * feat[.,c,.,.], grad_bias = sum grad_logits
 *                      (deterministic two-level reduction) -- replaces the convolution backward's three kernels.
 * feat: device, contiguous NCHW (B, C, H, W), PIL_F32 or PIL_BF16, 1 <= C <= 128; weight / bias / grad_weight /
 * grad_bias: device fp32 (bias, grad_bias may be NULL).  Both kernels are bound by the feature traffic (C * 4 bytes
 * per pixel each way); the fusion saves the logits / probability round trips and four launches.
 */
size_t pil_tail_workspace_bytes(int64_t C);
int pil_tail_forward(const void* feat, int feat_dtype, const float* weight, const float* bias, const void* t, int t_dtype,
                     int64_t B, int64_t C, int64_t H, int64_t W, int x_kind, const PilParams* p,
                     float* logits_out, double* sums, void* workspace, size_t workspace_bytes,
                     const struct PilExchange* ex /* may be NULL */, void* stream);
int pil_tail_backward(const void* feat, int feat_dtype, const float* weight, const float* grad_logits,
                      int64_t B, int64_t C, int64_t H, int64_t W, void* grad_feat, float* grad_weight, float* grad_bias,
                      void* workspace, size_t workspace_bytes, void* stream);

/*
 * Parameter sweeps (BASELINE config 4; the S2/S3 sensitivity grids of run_ablation.py:159-224 evaluated as
 * ONE batched loss evaluation).  lap(u), g = u(1-u), h = g*u, |grad u|^2 and the Dice/BCE sums do not
 * depend on D, a, eps or the weights, and r = D*lap + h - a*g, so the loss for ANY setting of the knobs is a
 * closed form in 13 parameter-independent sums.  pil_forward_moments makes one pass over the maps (same
 * fused kernel as pil_forward, 8 B/px) and pil_sweep_finalize turns the moments into n_params loss
 * reports (a one-block kernel; no further pass over the maps).
 *
 * moments (device, PIL_NMOMENTS doubles) for THIS shard -- all-reduce (SUM) across ranks before finalising:
 *   [0] I  [1] P  [2] T  [3] sum BCE terms  [4] sum lap^2  [5] sum gx^2+gy^2  [6] sum g^2  [7] n_invalid
 *   [8] sum lap*h  [9] sum lap*g  [10] sum h^2  [11] sum h*g  [12] n_pixels  [13..15] 0
 * params: HOST array of n_params settings (validated like pil_validate_params); loss_out: device,
 * n_params x PIL_NOUT floats, row k laid out like pil_finalize's loss_out for params[k].
 */
int pil_forward_moments(const void* x, const void* t, int64_t B, int64_t H, int64_t W,
                        int x_dtype, int t_dtype, int x_kind,
                        double* moments, void* workspace, size_t workspace_bytes, void* stream);
int pil_sweep_finalize(const double* moments, int64_t n_global, const PilParams* params, int n_params,
                       float* loss_out, void* stream);
/* The sweep over a batch sharded across ranks, without a collective call: the last block of the moments kernel stores
 * the shard's 16 sums into every rank's mailbox (two 8-double vectors, phases 0 and 1 of the PilExchange below);
 * pil_sweep_finalize_xchg waits for all ranks' vectors in the LOCAL mailbox, adds them in rank order and writes the
 * n_params (<= 32) loss reports of the GLOBAL batch, identical on every rank (moments_out: the global moments, may be
 * NULL).  Same epoch discipline as the training step: one epoch per sweep step; with PIL_XCHG_DEVICE_EPOCH the
 * finalize kernel completes the step (it advances the mailbox counter), so every pil_forward_moments_xchg must be
 * followed by its pil_sweep_finalize_xchg. */
int pil_forward_moments_xchg(const void* x, const void* t, int64_t B, int64_t H, int64_t W,
                             int x_dtype, int t_dtype, int x_kind,
                             double* moments, void* workspace, size_t workspace_bytes,
                             const struct PilExchange* ex, void* stream);
int pil_sweep_finalize_xchg(const struct PilExchange* ex, int64_t n_global, const PilParams* params, int n_params,
                            float* loss_out, double* moments_out, void* stream);

/*
 * Data-parallel training step over PEER MEMORY (one process per GPU of one NVLink/NVSwitch node).
 * The batch shards by whole images (every image has its own mirror boundary, src/pde.py:67, so no halo
 * exchange); the only coupling is the batch-global reductions of src/loss.py:134-141 and
 * src/pde.py:143,:210.  Instead of an NCCL all-reduce between the two kernels, the kernels exchange
 * their 8-double vectors themselves: the last block of pil_forward_pointwise_xchg stores the shard's
 * sums into every rank's mailbox (remote 8-byte stores over NVLink, each carrying the step's tag), every block of
 * pil_backward_accumulate_xchg waits on its LOCAL copy of the flags and adds the vectors in rank order
 * (bit-identical global sums on all ranks), and its last block swaps the stencil sums the same way
 * and finalises the GLOBAL loss report.  Two launches per step, no collective call, no host sync.
 *
 * Set-up (host, once): every rank calls pil_exchange_alloc, the ranks exchange the 64-byte IPC
 * handles out of band (torch.distributed / MPI / a file), and open each peer's with
 * pil_exchange_open; mailbox[rank] is the rank's own pointer.  Per step all ranks pass the SAME
 * `epoch`, incremented by one every step, and run the two calls in this order on one stream.
 * Every rank must own its GPU: the kernels wait on flags written by kernels of other ranks.
 * A wait that exceeds PIL_XCHG_TIMEOUT_MS (default 20000) gives up: the loss report becomes NaN, the rank's
 * gradient is written as ZEROS (a rank that lost its peers must not poison the weights) and the mailbox
 * status word (pil_exchange_status) is set.
 */
#define PIL_MAX_RANKS 8
#define PIL_IPC_HANDLE_BYTES 64
#define PIL_XCHG_DEFER_FINALIZE 1u /* the backward only pushes its stencil sums; pil_exchange_finalize assembles the loss */
/* The step tag comes from a counter in the rank's own mailbox instead of `epoch`: the backward's last block
 * increments it when the step is complete.  The two launches of a step then take identical arguments every
 * step, so they can be captured once in a CUDA graph (pil_step_graph_*).  All ranks must use the same mode for
 * the lifetime of a mailbox, and pil_forward_pointwise_xchg must always be followed by its
 * pil_backward_accumulate_xchg.  Not combinable with PIL_XCHG_DEFER_FINALIZE. */
#define PIL_XCHG_DEVICE_EPOCH 2u
typedef struct PilExchange {
    int32_t rank, world;
    uint64_t epoch;
    uint32_t flags, reserved;
    void* mailbox[PIL_MAX_RANKS]; /* device pointers valid in THIS process: own mailbox and the opened peers */
} PilExchange;
size_t pil_exchange_bytes(void);
int pil_exchange_alloc(void** mailbox, void* ipc_handle_out /* PIL_IPC_HANDLE_BYTES, may be NULL */);
int pil_exchange_open(const void* ipc_handle, void** peer_mailbox);
int pil_exchange_close(void* peer_mailbox);
int pil_exchange_free(void* mailbox);
int pil_exchange_status(const void* mailbox, int* status_out, void* stream); /* host sync; 0 = no timeout so far */
int pil_forward_pointwise_xchg(const void* x, const void* t, int64_t B, int64_t H, int64_t W,
                               int x_dtype, int t_dtype, int x_kind, const PilParams* p,
                               double* sums /* this shard's */, void* workspace, size_t workspace_bytes,
                               const PilExchange* ex, void* stream);
/* grad of THIS shard from the GLOBAL sums; stencil_sums: this shard's; loss_out (PIL_NOUT floats) and
 * total_sums (PIL_NSUMS doubles, may be NULL): the GLOBAL batch, identical on every rank. */
int pil_backward_accumulate_xchg(const void* x, const void* t, void* grad, int64_t B, int64_t H, int64_t W,
                                 int x_dtype, int t_dtype, int x_kind, const PilParams* p,
                                 const PilExchange* ex, int64_t n_global, const float* upstream, float grad_scale,
                                 double* stencil_sums, float* loss_out, double* total_sums,
                                 void* workspace, size_t workspace_bytes, void* stream);
/* Push a sums vector (device, PIL_NSUMS doubles) of this rank into every rank's mailbox, as the last block of
 * pil_forward_pointwise_xchg does: for callers that assemble the shard's pointwise sums (phase 0) from several
 * launches, e.g. chunk by chunk while the maps are still arriving from the host (pil_session_run_xchg). */
int pil_exchange_push(const PilExchange* ex, int phase, const double* sums, void* stream);
/* With PIL_XCHG_DEFER_FINALIZE the backward kernel does not wait for the other ranks' stencil sums;
 * this one-thread kernel does, later on the stream (e.g. after the optimizer step was enqueued).
 * It is also how several ranks are emulated on ONE GPU in the tests: there a kernel must never wait
 * for a flag that a LATER launch writes, so all backwards are enqueued first, then the finalizes. */
int pil_exchange_finalize(const PilExchange* ex, int64_t n_global, const PilParams* p,
                          float* loss_out, double* total_sums, void* stream);

/*
 * One training-step evaluation as ONE CUDA-graph launch.  For small maps -- the reference's real batches are
 * 8 x 1 x 128 x 128 (src/dataset.py:18), and a strong-scaled data-parallel batch leaves each GPU a small shard --
 * a step costs what the host spends on two C-ABI calls and two launches, not what the kernels take.
 * pil_step_graph_create runs the step once on `stream` (a real step; data parallel: on every rank, creation is
 * collective), then captures pil_forward_pointwise[_xchg] -> pil_backward_accumulate[_xchg] with their
 * programmatic-dependent-launch edge; pil_step_graph_launch replays the pair on any stream of the same device.
 * The graph is BOUND to the buffers it was created with (x, t, grad, sums, loss_out, stencil_sums, workspace,
 * upstream, total_sums): refill x / t in place between steps.  ex (may be NULL) must carry
 * PIL_XCHG_DEVICE_EPOCH; its mailbox pointers must stay mapped while the graph lives.  total_sums: the complete
 * sums vector (single shard: may alias sums; data parallel: the GLOBAL sums; may be NULL).
 */
typedef struct PilStepGraph PilStepGraph;
int pil_step_graph_create(PilStepGraph** out, const void* x, const void* t, void* grad, int64_t B, int64_t H, int64_t W,
                          int x_dtype, int t_dtype, int x_kind, const PilParams* p,
                          double* sums, float* loss_out, double* stencil_sums, void* workspace, size_t workspace_bytes,
                          const struct PilExchange* ex, int64_t n_global, const float* upstream, float grad_scale,
                          double* total_sums, void* stream);
int pil_step_graph_launch(PilStepGraph* g, void* stream);
int pil_step_graph_destroy(PilStepGraph* g);
/* The sweep step (pil_forward_moments[_xchg] -> pil_sweep_finalize[_xchg]) as one graph launch: a sharded sweep leaves
 * each GPU a pass of a few tens of microseconds, less than the host needs for two calls.  Same rules as above (bound to
 * x, t, moments, workspace, loss_out, moments_out; ex NULL or with PIL_XCHG_DEVICE_EPOCH; creation runs one real sweep
 * step and is collective).  params (HOST, n_params settings; <= 32 with ex) are copied into the graph.  Launch and
 * destroy with pil_step_graph_launch / pil_step_graph_destroy. */
int pil_sweep_graph_create(PilStepGraph** out, const void* x, const void* t, int64_t B, int64_t H, int64_t W,
                           int x_dtype, int t_dtype, int x_kind, double* moments, void* workspace, size_t workspace_bytes,
                           const struct PilExchange* ex, int64_t n_global, const PilParams* params, int n_params,
                           float* loss_out, double* moments_out, void* stream);

/*
 * The PDERegularization operators a caller can use on their own (fp32 maps):
 *   pil_laplacian            compute_laplacian              src/pde.py:49-79
 *   pil_laplacian_adjoint    its autograd backward (transpose of the reflect-Laplacian)
 *   pil_reaction             reaction_term                  src/pde.py:81-99
 *   pil_grad_mag_sq          compute_gradient_magnitude     src/pde.py:147-178
 *   pil_grad_mag_sq_backward its autograd backward: out = J^T (g), J the Jacobian at u
 */
int pil_laplacian(const float* u, float* out, int64_t B, int64_t H, int64_t W, void* stream);
int pil_laplacian_adjoint(const float* g, float* out, int64_t B, int64_t H, int64_t W, void* stream);
int pil_reaction(const float* u, float* out, int64_t n, double reaction_threshold, void* stream);
int pil_grad_mag_sq(const float* u, float* out, int64_t B, int64_t H, int64_t W, void* stream);
int pil_grad_mag_sq_backward(const float* u, const float* g, float* out, int64_t B, int64_t H, int64_t W,
                             void* stream);

/*
 * Host-buffer session: the end-to-end call for a caller whose maps live in HOST memory (pinned for
 * full speed).  One pil_session_run = H2D of x and t (chunked, overlapped with the forward kernel),
 * forward, backward (chunked, overlapped with the D2H of the gradient), D2H of the loss report.
 * It is what `loss = criterion(outputs, masks); loss.backward()` (src/train.py:117,:163) is to a
 * host-side caller.  grad_host may be NULL (loss only).  loss_out_host: PIL_NOUT floats.
 */
typedef struct PilSession PilSession;
int pil_session_create(PilSession** out, int device, int64_t max_B, int64_t H, int64_t W,
                       int x_dtype, int t_dtype);
int pil_session_run(PilSession* s, const void* x_host, const void* t_host, void* grad_host, int64_t B,
                    int x_kind, const PilParams* p, float* loss_out_host);
/* flags for pil_session_run_ex.  PIL_SESSION_GRAD_ON_DEVICE (with grad_host == NULL): forward AND backward
 * run, the gradient stays in the session's device buffer (pil_session_grad_ptr; valid until the next
 * run) the way a training step consumes it, and only the loss report crosses back to the host. */
#define PIL_SESSION_GRAD_ON_DEVICE 1
int pil_session_run_ex(PilSession* s, const void* x_host, const void* t_host, void* grad_host, int64_t B,
                       int x_kind, const PilParams* p, float* loss_out_host, int flags);
/* Data-parallel form (one process per GPU, one session per rank): x_host / t_host hold THIS rank's B images of
 * the global batch.  H2D in chunks overlapped with the pointwise forward, push of the shard's sums to every rank
 * (pil_exchange_push), then ONE backward over the shard fed from the mailbox (pil_backward_accumulate_xchg): the
 * gradient of the GLOBAL loss w.r.t. this rank's maps stays on the device (pil_session_grad_ptr), and
 * loss_out_host receives the GLOBAL loss report, identical on every rank.  n_global: pixels of the global batch
 * (<= 0: taken from the exchanged sums); grad_scale: world size under DDP-style gradient averaging. */
int pil_session_run_xchg(PilSession* s, const void* x_host, const void* t_host, int64_t B, int x_kind,
                         const PilParams* p, const struct PilExchange* ex, int64_t n_global, float grad_scale,
                         float* loss_out_host);
void* pil_session_grad_ptr(PilSession* s);
int pil_session_destroy(PilSession* s);

/* Introspection for tests/benchmarks: how the last launch on this thread was tiled. */
typedef struct PilLaunchInfo {
    int32_t fwd_blocks, fwd_threads, fwd_rows_per_segment, fwd_aligned;
    int32_t bwd_blocks, bwd_threads, bwd_rows_per_segment, bwd_aligned;
    int64_t kernels_launched; /* running count of kernels this library launched in this process */
} PilLaunchInfo;
int pil_last_launch_info(PilLaunchInfo* out);

/* Tuning override for benchmarks (0 = automatic): rows per segment for fwd / bwd. */
int pil_set_tuning(int fwd_rows_per_segment, int bwd_rows_per_segment);
/* How the aligned backward kernel stages its rows in shared memory: 1 = TMA boxes (cp.async.bulk.tensor, one
 * elected lane per warp, mbarrier completion), 0 = one cp.async per lane per row, negative = default
 * (PIL_BWD_STAGE=tma|cpasync, else the library's choice).  PilLaunchInfo.bwd_aligned reports 2 for TMA. */
int pil_set_bwd_staging(int mode);
/* How many MB at the end of each fp32 map the pointwise forward asks L2 to keep (evict_last) for the backward
 * kernel, which starts there; the rest of its stream is evict_first.  0 = plain loads; negative = default
 * (PIL_L2_KEEP_MB or 12). */
int pil_set_l2_keep_mb(int mb);

#ifdef __cplusplus
}
#endif
#endif /* PIL_H_ */
