/*
 * mms_b200.h -- C ABI of libmms_b200.so: hand-written sm_100a kernels for the hot path of
 * 17LiQi/MultimodalSignal (CnnGruAttentionModel training step + resample/windowing).
 *
 * The reference has no FFI of its own (pure Python / torch.nn); the boundary
 * it offers is its Python API.  Every entry point below therefore replaces a *library call
 * site* of the reference, cited as file:line into /root/reference.  The Python host side
 * (the multimodalsignal_b200 package) binds these with ctypes and keeps the reference's class and
 * function signatures.
 *
 * Conventions
 *   - every function returns 0 on success or a negative MMS_E_* code; mms_last_error()
 *     returns a thread-local description of the most recent failure;
 *   - all pointers are DEVICE pointers unless the parameter name ends in _host;
 *   - nothing here allocates, frees or synchronises: workspaces are caller-provided (sizes
 *     from the *_workspace_bytes functions) and every launch goes to the cudaStream_t given
 *     (passed as void* so that this header needs no CUDA headers);
 *   - there is no CPU fallback: mms_init() fails unless the device is compute capability 10.x.
 *   - tensors are contiguous float32 unless stated otherwise.
 */
#ifndef MMS_B200_H
#define MMS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMS_OK 0
#define MMS_E_INVALID (-1)      /* bad argument / unsupported shape            */
#define MMS_E_CUDA (-2)         /* a CUDA runtime call or launch failed        */
#define MMS_E_ARCH (-3)         /* device is not sm_100 (no fallback exists)   */
#define MMS_E_WORKSPACE (-4)    /* workspace too small                         */

typedef void* mms_stream_t;     /* cudaStream_t */

int mms_version(void);
const char* mms_last_error(void);
/* Select `device`, verify it is sm_100 (B200) and raise the dynamic shared-memory limits of
 * the kernels that need it.  Must be called once per process before any other function. */
int mms_init(int device);

/* Diagnostics.  mms_launch_count(): kernels launched by this library since it was loaded
 * (launches recorded into a CUDA graph count once, at capture).  mms_profile_enable(1) makes
 * every subsequent eager launch record a CUDA-event pair on its stream (and clears earlier
 * records); mms_profile_report() synchronises the device and writes one line per kernel name,
 * "name launches total_ms", in first-launch order, into a host buffer. */
int64_t mms_launch_count(void);
int mms_profile_enable(int32_t on);
int mms_profile_report(char* buf_host, int64_t buf_bytes);
/* In-graph timeline (diagnostic; the one place the library allocates: a 16 KB device buffer on first enable).  While
 * enabled, every launch is bracketed by one-thread kernels that store %globaltimer, on the launch's own stream -- also
 * under CUDA-graph capture, where the stamps become graph nodes.  mms_timeline_report() synchronises the device and
 * writes one line per launch of the LAST execution, "name stream_index start_ns end_ns" (relative to the earliest
 * start), into a host buffer.  Enable before capturing / launching, report after; enabling clears earlier records. */
int mms_timeline_enable(int32_t on);
int mms_timeline_report(char* buf_host, int64_t buf_bytes);
/* Weight-gradient kernels normally run on two library-owned side streams (forked from and joined
 * back into the caller's stream, also under CUDA-graph capture).  0 serialises everything on the
 * caller's stream (used while per-kernel times are taken); also MMS_DISABLE_STREAMS=1. */
int mms_set_side_streams(int32_t on);
/* Integer run-time switches for A/B measurements and kernel selection.  mms_set_option("GRU_BWD_RING", 0) has the
 * effect of the environment variable MMS_GRU_BWD_RING=0 but can be changed between launches (the environment is read
 * once, at first use); mms_get_option returns the value in effect (`dflt` if neither was given; defaults are per call
 * site and listed in INTEGRATION.md).  Options never change
 * results beyond the documented tolerances; they select between kernels of this library. */
int mms_set_option(const char* name, int32_t value);
int32_t mms_get_option(const char* name, int32_t dflt);
int mms_clear_option(const char* name);     /* back to the built-in default of every call site */

/* ------------------------------------------------------------------------------------------
 * Model description (reference models.py:39-40 constructor arguments + call-time facts).
 * Conv widths 16 / k7 / k5 and the 64-wide classifier hidden layer are hard-coded in the
 * reference (models.py:46,50,67) and therefore here.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t batch;          /* B                                                        */
    int32_t in_channels;    /* C  = len(CHANNELS_TO_USE), main.py:47,116                */
    int32_t seq_len;        /* T  = window samples (3840 @64 Hz, 7680 @128 Hz)          */
    int32_t num_classes;    /* 2 (stress_binary) or 3 (ternary), <= 8                   */
    int32_t cnn_out;        /* cnn_out_channels, models.py:39 (32)                      */
    int32_t hidden;         /* gru_hidden_size (64 or 32)                               */
    int32_t layers;         /* gru_num_layers (>= 1)                                    */
    int32_t training;       /* 1: batch-stat BN + dropout (model.train()), 0: eval      */
    int32_t attention;      /* 1: ChannelAttention gate, 0: cnn_gru baseline*/
    int32_t need_grad;      /* 1: keep activations for mms_cnngru_backward              */
    float dropout_p;        /* models.py:40; applied between GRU layers and in the head */
    uint64_t rng_seed;      /* dropout stream seed                                      */
    uint64_t rng_offset;    /* dropout stream offset used when rng_offset_dev == NULL   */
    const int64_t* rng_offset_dev; /* optional device counter read by the kernels
                                      (the fused train step passes its Adam step count) */
    int32_t global_batch;   /* 0 or == batch: single GPU.  > batch: intra-fold data parallelism -- BatchNorm
                               statistics are counted over the global batch (SyncBN) and the BN affine
                               gradients are scaled by batch/global_batch so that a sum all-reduce is exact */
} mms_cnngru_desc;

/* Flat parameter buffer.  The model's parameters live in ONE float32 buffer so that Adam and
 * the gradient all-reduce are single launches.  Segment order (each segment 16-byte aligned):
 *   0 ca_w1   [C/4, C]      channel_attention.fc.0.weight      models.py:18
 *   1 ca_w2   [C, C/4]      channel_attention.fc.2.weight      models.py:20
 *   2 conv1_w [16, C, 7]    cnn_encoder.0.weight               models.py:46
 *   3 bn1_g [16]  4 bn1_b [16]                                  models.py:47
 *   5 conv2_w [O, 16, 5]    cnn_encoder.4.weight               models.py:50
 *   6 bn2_g [O]   7 bn2_b [O]                                   models.py:51
 *   then for each GRU layer l (models.py:56-63), both directions adjacent (forward first):
 *   8+4l w_ih [2][3H, I_l]   9+4l w_hh [2][3H, H]   10+4l b_ih [2][3H]   11+4l b_hh [2][3H]
 *   then fc0_w [64, 2H], fc0_b [64], fc3_w [nc, 64], fc3_b [nc]  models.py:66-71
 * offsets/sizes are in floats.  Returns the number of segments (or < 0). */
#define MMS_MAX_SEGMENTS 64
int mms_cnngru_param_layout(const mms_cnngru_desc* d, int64_t* offsets_host, int64_t* sizes_host,
                            int32_t max_segments, int64_t* total_floats_host);

/* BatchNorm buffers: bn_buffers = [running_mean1[16] | running_var1[16] | running_mean2[O] |
 * running_var2[O]] float32, num_batches_tracked = int64[2] (state_dict entries
 * cnn_encoder.{1,5}.running_mean/.running_var/.num_batches_tracked). */

int64_t mms_cnngru_workspace_bytes(const mms_cnngru_desc* d);

/* models.py:73-81 forward.  x [B,C,T] -> logits [B,nc].  In training mode the workspace keeps
 * every activation the backward needs and the BN running statistics are updated. */
int mms_cnngru_forward(const mms_cnngru_desc* d, const float* x, const float* params,
                       float* bn_buffers, int64_t* num_batches_tracked, void* workspace,
                       float* logits, mms_stream_t stream);

/* trainer.py:148 loss.backward() from dlogits [B,nc]: adds every parameter gradient into
 * grads (same flat layout as params; the caller zeroes it, trainer.py:144 zero_grad).
 * dx may be NULL (inputs do not require grad in the reference, trainer.py:140). */
int mms_cnngru_backward(const mms_cnngru_desc* d, const float* x, const float* params,
                        const float* bn_buffers, void* workspace, const float* dlogits,
                        float* grads, float* dx, mms_stream_t stream);

/* Data-parallel form of the two calls above (intra-fold DP).  `phases` is a bit mask:
 *   forward : 1 = gate + conv1 (+BN1 sums) | 2 = BN1/pool1 + conv2 (+BN2 sums) | 4 = the rest
 *   backward: 1 = head..GRU..stage-2 pool/ReLU backward (+BN2 reductions) | 2 = stage-2 BN apply, conv2
 *             gradients, stage-1 pool/ReLU backward (+BN1 reductions) | 4 = the rest
 * Between the phases the caller sum-all-reduces the float64 vectors whose byte offsets inside the
 * workspace mms_cnngru_sync_offsets() reports: [0] BN1 sums, [1] BN2 sums (forward), [2] BN1
 * reductions, [3] BN2 reductions (backward); counts are in doubles.  desc.global_batch must be set. */
int mms_cnngru_forward_phase(const mms_cnngru_desc* d, int32_t phases, const float* x, const float* params,
                             float* bn_buffers, int64_t* num_batches_tracked, void* workspace, float* logits,
                             mms_stream_t stream);
int mms_cnngru_backward_phase(const mms_cnngru_desc* d, int32_t phases, const float* x, const float* params,
                              const float* bn_buffers, void* workspace, const float* dlogits, float* grads,
                              mms_stream_t stream);
int mms_cnngru_sync_offsets(const mms_cnngru_desc* d, int64_t* byte_offsets_host, int64_t* counts_host);
/* Cross entropy of a rank's share of a global batch: loss_out = sum_local(...)/global_batch (the per-rank
 * values add up to the global mean), dlogits = (softmax - onehot)/global_batch. */
int mms_cross_entropy_partial(const float* logits, const int64_t* labels, int32_t batch, int32_t num_classes,
                              int32_t global_batch, float* loss_out, float* dlogits, double* loss_sum_accum,
                              mms_stream_t stream);

/* The two exchange steps of intra-fold data parallelism WITHOUT a communication library on the data path:
 * the peers' buffers (symmetric memory mapped into this process, e.g. torch.distributed._symmetric_memory) are read
 * directly over NVLink and the ranks synchronise through flags in each other's signal pads.
 *   bufs_host / grads_host : HOST array of `world` device pointers, entry p = rank p's copy of the buffer;
 *   signals_host           : HOST array of `world` device pointers to the ranks' signal pads (uint32 words, zero at start);
 *                            a call uses words [signal_base, signal_base + 2*world) of every pad;
 *   epoch_dev              : local uint32 counter (zero at start), advanced by the call.
 * Every rank must make the same sequence of calls.  Nothing here blocks the host.
 * mms_peer_allreduce_f64: in-place sum of `count` (<= 256) doubles -- the SyncBN (sum, sum of squares) / (sum dy, sum dy*xhat)
 * vectors whose workspace offsets mms_cnngru_sync_offsets() reports.
 * mms_peer_allreduce_adam: flat gradient all-reduce FUSED with mms_adam_flat_step (trainer.py:148-149): the gradients of
 * all ranks are summed on the fly in rank order (bit-identical parameters on every rank) and consumed by the Adam update;
 * scratch2_dev: two zero-initialised local uint32.
 * Every wait inside these kernels is bounded (MMS_PEER_TIMEOUT_MS, default 2000): if a peer never signals -- a rank died,
 * made a different number of steps, or raised between two phases -- the kernel gives up, finishes with a meaningless sum and
 * counts the event.  mms_peer_status() returns that count for the current device and clears it (it synchronises the device);
 * a caller checks it wherever it reads results back.  The reference has no counterpart (its training is single-device,
 * trainer.py:57-58); this is the failure detection of the data-parallel extension (BASELINE.json configs[4]). */
int mms_peer_status(uint32_t* timeouts_host);
int mms_peer_allreduce_f64(const void* const* bufs_host, void* const* signals_host, int32_t world, int32_t rank,
                           int32_t signal_base, int32_t count, uint32_t* epoch_dev, mms_stream_t stream);
int mms_peer_allreduce_adam(float* params, const void* const* grads_host, void* const* signals_host, int32_t world,
                            int32_t rank, int32_t signal_base, float* exp_avg, float* exp_avg_sq, int64_t n,
                            const float* lr_dev, float beta1, float beta2, float eps, float weight_decay,
                            int64_t* step_dev, uint32_t* epoch_dev, uint32_t* scratch2_dev, mms_stream_t stream);

/* trainer.py:69,147 CrossEntropyLoss() (mean): loss_out[0] = mean_b(lse - logit[y]),
 * dlogits = (softmax - onehot)/B (NULL to skip).  If loss_sum_accum != NULL it receives
 * += loss * B in float64 (trainer.py:152 accumulates loss.item()*batch_size on the host;
 * this keeps the sum on the device so that a step needs no sync). */
int mms_cross_entropy(const float* logits, const int64_t* labels, int32_t batch, int32_t num_classes,
                      float* loss_out, float* dlogits, double* loss_sum_accum, mms_stream_t stream);

/* trainer.py:68,149 torch.optim.Adam (coupled L2 weight decay) over the flat buffer.
 * step_dev holds the number of steps already taken; the kernel uses step_dev+1 for the bias
 * corrections and increments it.  lr_dev is a device scalar (ReduceLROnPlateau writes it,
 * trainer.py:72-77,160).  scratch_dev: one int32, zero-initialised. */
int mms_adam_flat_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                       int64_t n, const float* lr_dev, float beta1, float beta2, float eps,
                       float weight_decay, int64_t* step_dev, int32_t* scratch_dev, mms_stream_t stream);

/* trainer.py:130,140-142 -- one batch of DataLoader(dataset, batch_size, shuffle=True) (main.py:111) assembled on
 * the device: out_x[b, :] = data[perm[cursor + b], :], out_y[b] = labels[perm[cursor + b]] for b < batch.
 * data: float32 [n_rows, row_floats] (the normalised [N, C*W] window array), labels int64 [n_rows], perm: int64
 * permutation (NULL = identity), *cursor_dev: int64 position inside perm (NULL = 0).  With advance != 0 the kernel
 * moves *cursor_dev forward by `batch` when it is done, so a captured CUDA graph (gather + train step) can be
 * replayed once per batch with no host work; scratch_dev: one zero-initialised int32. */
int mms_batch_gather(const float* data, const int64_t* labels, const int64_t* perm, int64_t* cursor_dev,
                     int64_t n_rows, int64_t row_floats, int32_t batch, float* out_x, int64_t* out_y,
                     int32_t advance, int32_t* scratch_dev, mms_stream_t stream);

/* trainer.py:217-228 -- the per-batch bookkeeping of Trainer.evaluate: *loss_sum += sum_b(lse_b - logit[b, y_b])
 * (float64; the reference adds loss.item() * batch), preds_out[b] = argmax_c logits[b, c] (first maximum, as
 * torch.argmax of the softmax), confusion[y_b * nc + pred_b] += 1 (int64 [nc, nc], rows = true class).
 * accuracy_score and the weighted f1_score of trainer.py:229-230 are functions of that matrix.
 * preds_out, confusion and loss_sum may each be NULL. */
int mms_eval_accumulate(const float* logits, const int64_t* labels, int32_t batch, int32_t num_classes,
                        int64_t* preds_out, int64_t* confusion, double* loss_sum, mms_stream_t stream);

/* One whole training step, trainer.py:144-149: zero_grad, forward, CE, backward, Adam.
 * (With world_size > 1 the host calls forward/backward/adam separately around its NCCL
 * all-reduce instead.) */
int mms_cnngru_train_step(const mms_cnngru_desc* d, const float* x, const int64_t* labels,
                          float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                          float* bn_buffers, int64_t* num_batches_tracked, void* workspace,
                          float* logits, float* loss_out, double* loss_sum_accum,
                          const float* lr_dev, float beta1, float beta2, float eps, float weight_decay,
                          int64_t* step_dev, int32_t* scratch_dev, mms_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Operator-level entry points (the same kernels the composed calls launch), for unit parity
 * tests and for callers that use one layer on its own.
 * ---------------------------------------------------------------------------------------- */

/* models.py:24-31 ChannelAttention.forward.  hidden = C/4 may be 0 (gate == 0.5).
 * mean_out/gate_out [B,C]; y may be NULL (gate only; the model folds the gate into conv1). */
int mms_chan_attn_fwd(const float* x, const float* w1, const float* w2, int32_t B, int32_t C, int32_t T,
                      float* mean_out, float* gate_out, float* y, mms_stream_t stream);
/* Backward of the above: dy [B,C,T] -> dx [B,C,T] (may be NULL), dw1, dw2 (+=). scratch: 4*B*C floats. */
int mms_chan_attn_bwd(const float* x, const float* dy, const float* w1, const float* w2,
                      const float* mean, const float* gate, int32_t B, int32_t C, int32_t T,
                      float* dx, float* dw1, float* dw2, float* scratch, mms_stream_t stream);

/* models.py:46 / :50  Conv1d(bias=False).  which = 1: k7 s2 p3 (C_in = any <= 16, C_out 16);
 * which = 2: k5 s2 p2 (C_in 16, C_out = cnn_out).  gate [B,C_in] optional (x*gate fused).
 * stats (float64 [2][C_out]: sum, sum of squares over (B,L_out)) optional, accumulated (+=). */
int mms_conv1d_fwd(int32_t which, const float* x, const float* w, const float* gate, int32_t B,
                   int32_t c_in, int32_t c_out, int32_t l_in, float* y, double* stats, mms_stream_t stream);
/* The same convolution as an implicit GEMM on the tensor cores: TMA-staged input tile (the zero padding is TMA's
 * out-of-bounds fill), im2col + 3xTF32 operand expansion in shared memory, tcgen05.mma.kind::tf32 with the accumulator in
 * TMEM.  Same arguments and results (fp32-class accuracy); needs l_in % 4 == 0, l_in >= 64, x 16-byte aligned and
 * C_out <= 32.  mms_conv1d_fwd uses it when MMS_CONV_TC=1 (measured slower than the SIMT kernel at K = 42 / 80). */
int mms_conv1d_fwd_tc(int32_t which, const float* x, const float* w, const float* gate, int32_t B,
                      int32_t c_in, int32_t c_out, int32_t l_in, float* y, double* stats, mms_stream_t stream);
/* dgrad: dy [B,C_out,L_out] -> dx [B,C_in,L_in] (may be NULL) and, if xdot != NULL,
 * dgate[b,c] += sum_t dx[b,c,t]*xdot[b,c,t] (the only part of conv1's dgrad training needs). */
int mms_conv1d_dgrad(int32_t which, const float* dy, const float* w, int32_t B, int32_t c_in, int32_t c_out,
                     int32_t l_in, float* dx, const float* xdot, float* dgate, mms_stream_t stream);
/* wgrad: dw[o,c,k] += sum_{b,l} dy[b,o,l] * gate[b,c] * x[b,c,s*l+k-p]. */
int mms_conv1d_wgrad(int32_t which, const float* x, const float* dy, const float* gate, int32_t B,
                     int32_t c_in, int32_t c_out, int32_t l_in, float* dw, mms_stream_t stream);

/* models.py:47-49 / :51-53  BatchNorm1d + ReLU + MaxPool1d(3,2,1) fused.
 * training: batch statistics from `stats` (as produced by mms_conv1d_fwd), running buffers
 * updated (momentum 0.1, unbiased variance), nbt incremented; eval: running buffers.
 * time_major = 1 writes out[b, l, c] (models.py:77 permute(0,2,1) for free). */
int mms_bn_relu_pool_fwd(const float* y, const double* stats, const float* gamma, const float* beta,
                         float* running_mean, float* running_var, int64_t* nbt, int32_t B, int32_t C,
                         int32_t l_in, int32_t training, int32_t time_major, float* out, mms_stream_t stream);
/* Backward: dout (layout per time_major) -> dy [B,C,L_in] (in-place capable scratch), dgamma/dbeta (+=).
 * red: float64 [2][C] zero-initialised scratch. */
int mms_bn_relu_pool_bwd(const float* y, const double* stats, const float* gamma, const float* beta,
                         const float* running_mean, const float* running_var, const float* dout,
                         int32_t B, int32_t C, int32_t l_in, int32_t training, int32_t time_major,
                         float* dy, float* dgamma, float* dbeta, double* red, mms_stream_t stream);

/* GEMMs used for the GRU input projections and their gradients (models.py:78, the
 * x @ W_ih^T + b_ih part of nn.GRU):
 *   nt : C[m,n]  = sum_k A[m*lda+k] * W[n*ldw+k] + bias[n]
 *   nn : C[m,n] (+)= sum_k A[m*lda+k] * W[k*ldw+n]
 *   tn : C[i,j] += sum_m A[m*lda+acol(i)] * Bm[row(m)*ldb+j], bias_grad[i] += sum_m A[m*lda+acol(i)]
 *        acol(i) = i < a_split ? i : i + a_skip;  row(m) = m + shift within each block of
 *        `seq` rows (rows shifted outside their block read as zero) -- this is h_{t-1}. */
/* Tensor-core version of the nt form: TMA-staged operands, tcgen05.mma kind::tf32 with the 3xTF32
 * operand split (fp32-level accuracy), accumulators in TMEM.  C[m,n] (+)= sum_k A[m,k] W[n,k] + bias[n].
 * A and W must be 16-byte aligned with lda, ldw multiples of 4. */
int mms_tc_gemm_nt(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                   float* C, int64_t ldc, int32_t M, int32_t N, int32_t K, int32_t accumulate, mms_stream_t stream);
/* Tensor-core version of the tn form below (same arguments and meaning as mms_gemm_tn_acc): both row-major operands
 * are consumed MN-major by tcgen05.mma.kind::tf32 (3xTF32 split), split-K over the grid, partial tiles added to C
 * with red.global.add; the bias gradient is the product with an implicit column of ones.  Requires 16-byte aligned
 * operands, lda/ldb/ldc % 4 == 0, N1 % 32 == 0 (<= 256), N2 % 32 == 0 (<= 224, 0 = bias gradient only),
 * a_split % 32 == 0, a_skip % 32 == 0, M >= 128. */
int mms_tc_gemm_tn(const float* A, int64_t lda, int32_t a_split, int32_t a_skip, const float* Bm, int64_t ldb,
                   int32_t shift, int32_t seq, float* C, int64_t ldc, float* bias_grad,
                   int32_t M, int32_t N1, int32_t N2, mms_stream_t stream);
/* Several tensor-core TN products (2..4, each within the limits of mms_tc_gemm_tn) in ONE launch: the weight-gradient
 * products of a GRU layer (reference models.py:56-63 backward: dW_ih, dW_hh, db_ih, db_hh of both directions) reduce over
 * the same B*L rows, so one grid of ~148 CTAs covers all of them.  Used by the model when MMS_TN_BATCH=1. */
typedef struct {
    const float* A; int64_t lda; int32_t a_split, a_skip;
    const float* Bm; int64_t ldb; int32_t shift, seq;
    float* C; int64_t ldc; float* bias_grad;
    int32_t M, N1, N2;
} mms_tn_call;
int mms_tc_gemm_tn_batch(const mms_tn_call* calls_host, int32_t n, mms_stream_t stream);
int mms_gemm_nt_bias(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                     float* C, int64_t ldc, int32_t M, int32_t N, int32_t K, mms_stream_t stream);
int mms_gemm_nn(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc,
                int32_t M, int32_t N, int32_t K, int32_t accumulate, mms_stream_t stream);
/* Few-row products (M <= 256, K <= 256, K % 4 == 0, 16-byte aligned operands with strides that are multiples of 4): the B
 * rows of the top GRU layer's single reverse step (models.py:79 keeps only outputs[:, -1, :]) -- its input
 * projection in the forward pass and the input gradient of that step in the backward pass.
 *   C[m,n] = sum_k a(m,k) W(n,k) (+ bias[n]);  W(n,k) = W[n*ldw + k] (w_kmajor != 0) or W[k*ldw + n]
 *   drop_mode 1: a(m,k) = A[m,k] * m(drop_base + m*drop_row_stride + k)  (models.py:62 inter-layer dropout on the operand)
 *   drop_mode 2: C[m,n] *= m(drop_base + m*drop_row_stride + n)          (the same multipliers on the gradient)
 * with the multipliers of mms_dropout_apply. */
int mms_gemm_skinny(const float* A, int64_t lda, const float* W, int64_t ldw, int32_t w_kmajor, const float* bias,
                    float* C, int64_t ldc, int32_t M, int32_t N, int32_t K, int32_t drop_mode, int64_t drop_base,
                    int64_t drop_row_stride, float dropout_p, uint64_t rng_seed, uint64_t rng_offset,
                    const int64_t* rng_offset_dev, mms_stream_t stream);
int mms_gemm_tn_acc(const float* A, int64_t lda, int32_t a_split, int32_t a_skip, const float* Bm, int64_t ldb,
                    int32_t shift, int32_t seq, float* C, int64_t ldc, float* bias_grad,
                    int32_t M, int32_t N1, int32_t N2, mms_stream_t stream);

/* models.py:62 inter-layer GRU dropout as a streaming pass: out[i] = in[i] * m(base_id + i) with
 * m in {0, 1/(1-p)} drawn from the counter-based stream (seed, offset, element id); in == out is
 * allowed.  Forward and backward call it with the same ids, so they see the same mask. */
int mms_dropout_apply(const float* in, float* out, int64_t n, int64_t base_id, float dropout_p,
                      uint64_t rng_seed, uint64_t rng_offset, const int64_t* rng_offset_dev, mms_stream_t stream);

/* The GRU recurrence of one direction (models.py:56-63; gate order r,z,n; h0 = 0).
 * gi[(b*gi_bs + t*gi_ts) + 0..3H) are the input projections (incl. b_ih).  Runs `nsteps` steps
 * starting at t0 and moving by dt (+1 forward, -1 reverse).  Writes h to hs[b*hs_bs + t*hs_ts + j]
 * and, if stash != NULL, (r,z,n,W_hn h + b_hn) to
 * stash[(b*st_bs + t*st_ts) + 0..4H). */
typedef struct {
    const float* gi; int64_t gi_bs, gi_ts;
    const float* w_hh; const float* b_hh;
    float* hs; int64_t hs_bs, hs_ts;
    float* hs_drop; int64_t drop_base;   /* reserved, must be NULL / 0: use mms_dropout_apply */
    float* stash; int64_t st_bs, st_ts;
    int32_t t0, dt, nsteps;
} mms_gru_dir_fwd;
int mms_gru_recur_fwd(const mms_gru_dir_fwd* dirs_host, int32_t ndirs, int32_t B, int32_t H,
                      float dropout_p, uint64_t rng_seed, uint64_t rng_offset, const int64_t* rng_offset_dev,
                      mms_stream_t stream);

/* Reverse-time pass of one direction.  dout (optional) is the gradient w.r.t. the emitted h,
 * indexed like hs; dout_last [B, dl_ld] (optional) is added at the forward-order LAST step
 * only (or at the first one if dl_at_first); dh_head/W0 (optional): initial dh[b,k] += sum_i dh_head[b*64+i] * w0[i*w0_ld + w0_col + k].
 * drop_base / drop_mask are reserved and must be 0 (apply mms_dropout_apply to dout first).  Writes D[(b*d_bs + t*d_ts) + 0..4H) = (d r_pre, d z_pre, d n_pre, d q). */
typedef struct {
    const float* w_hh;
    const float* stash; int64_t st_bs, st_ts;
    const float* hs; int64_t hs_bs, hs_ts;
    const float* dout; int64_t do_bs, do_ts; int64_t drop_base; int32_t drop_mask;
    const float* dout_last; int64_t dl_ld; int32_t dl_at_first;  /* 1: add dout_last at the FIRST forward step instead */
    const float* dh_head; const float* w0; int64_t w0_ld; int32_t w0_col;
    float* D; int64_t d_bs, d_ts;
    int32_t t0, dt, nsteps;
} mms_gru_dir_bwd;
int mms_gru_recur_bwd(const mms_gru_dir_bwd* dirs_host, int32_t ndirs, int32_t B, int32_t H,
                      float dropout_p, uint64_t rng_seed, uint64_t rng_offset, const int64_t* rng_offset_dev,
                      mms_stream_t stream);

/* models.py:79-80 classifier head on last [B,2H]: Linear(2H,64)+ReLU+Dropout+Linear(64,nc).
 * hid_out [B,64] keeps the post-ReLU activations for the backward. */
int mms_head_fwd(const float* last, const float* w0, const float* b0, const float* w3, const float* b3,
                 int32_t B, int32_t H2, int32_t nc, float dropout_p, uint64_t rng_seed, uint64_t rng_offset,
                 const int64_t* rng_offset_dev, float* hid_out, float* logits, mms_stream_t stream);
int mms_head_bwd(const float* last, const float* hid, const float* dlogits, const float* w3,
                 int32_t B, int32_t H2, int32_t nc, float dropout_p, uint64_t rng_seed, uint64_t rng_offset,
                 const int64_t* rng_offset_dev, float* dhid, float* dw0, float* db0, float* dw3, float* db3,
                 mms_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Preprocess path (reference preprocess.py:70-75 resample_signal -> scipy.signal.resample,
 * preprocess.py:189-200 window stacking, dataset.py:37-48 normalisation).
 * ---------------------------------------------------------------------------------------- */

/* FFT resampling of `n_sig` real float64 signals of length n_in to length n_out
 * (scipy.signal.resample semantics, Fourier method, no window).  x [n_sig][n_in] ->
 * y [n_sig][n_out], both float64, rows contiguous. */
int64_t mms_resample_workspace_bytes(int64_t n_in, int64_t n_out, int32_t n_sig);
int mms_resample_f64(const double* x, int64_t n_in, int64_t n_out, int32_t n_sig, double* y,
                     void* workspace, int64_t workspace_bytes, mms_stream_t stream);

/* preprocess.py:189-200: out[w, i, c] = streams[c][starts[w] + i] for w < n_win, i < win, c < n_ch.
 * streams: HOST array of n_ch device pointers (float64 rows of length stream_len); out float64 [n_win, win, n_ch]
 * (out_f32 = 0) or float32 [n_win, n_ch, win] already permuted as dataset.py:63 does (out_f32 = 1,
 * in which case (x - shift[c]) * scale[c] is applied, log1p first where log_flag[c] != 0). */
int mms_window_gather(const double* const* streams, int32_t n_ch, int64_t stream_len,
                      const int64_t* starts, int32_t n_win, int32_t win, int32_t out_f32,
                      const double* shift, const double* scale, const int32_t* log_flag,
                      void* out, mms_stream_t stream);

/* dataset.py:37-48 statistics over the WINDOWED array without materialising it: for each
 * channel, sum and sum of squares (of log1p(x) where log_flag) over all (window, sample)
 * pairs, i.e. each stream sample weighted by the number of windows covering it.
 * sums float64 [n_ch][2], zero-initialised by the caller. */
int mms_window_stats(const double* const* streams, int32_t n_ch, int64_t stream_len,
                     const int64_t* starts, int32_t n_win, int32_t win, const int32_t* log_flag,
                     double* sums, mms_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MMS_B200_H */
