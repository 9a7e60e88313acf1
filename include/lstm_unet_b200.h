/*
 * lstm_unet_b200.h -- C-ABI of the B200-native ConvLSTM-UNet hot path (liblstm_unet_b200.so).
 *
 * The reference (arbellea/LSTM-UNet) has no FFI: its boundary for this path is the Python object protocol of
 * Networks.ULSTMnet2D as used by train2D.py / Inference2D.py.  Each entry point below names the reference
 * interface it replaces (file:line in the reference repository).  The Python mirror of that protocol
 * (lstm_unet_b200/Networks.py) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; lu_last_error() returns the message of the last
 *     failing call on this thread (the reference raises Python ValueError; the shim re-raises it).
 *   - all `dev` pointers are device pointers owned by the caller (torch.Tensor.data_ptr()); `stream` is a
 *     cudaStream_t passed as void* (0 = default stream).  Work is enqueued on `stream`, nothing synchronises
 *     unless stated.  One handle per GPU per model instance; not thread-safe (the reference calls the model
 *     from the main thread only).
 *   - there is NO CPU fallback: on a box without an sm_100 device every compute entry point fails.
 */
#ifndef LSTM_UNET_B200_H
#define LSTM_UNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LU_MAX_LEVELS 4
#define LU_MAX_PER_LEVEL 4

/* tensor-core operand precision: bf16 = throughput / training mode; bf16x3 = split-bf16, fp32-equivalent
 * (3x the MMA work); fp16 = 11-bit mantissa operands at the bf16 rate, inference handles only */
enum { LU_PREC_BF16 = 0, LU_PREC_BF16X3 = 1, LU_PREC_FP16 = 2 };
enum { LU_ENGINE_TCGEN05 = 0, LU_ENGINE_SIMT = 1 };   /* SIMT = on-GPU scalar mirror used to debug the TC path */
enum { LU_GATE_HARD_SIGMOID = 0, LU_GATE_SIGMOID = 1 };
enum { LU_AMODE_HALO = 0, LU_AMODE_DIRECT = 1 };      /* how activation tiles are staged in shared memory */
/* what the handle runs: the whole ULSTMnet2D, or one of its blocks on its own the way the reference's
 * DownBlock2D.unit_test / UpBlock2D.unit_test construct and call them (Networks.py:100-119,155-175) */
enum { LU_BLOCK_NET = 0, LU_BLOCK_DOWN = 1, LU_BLOCK_UP = 2 };

/* Architecture-as-data: mirrors the `net_kernel_params` dict (Params.py:49-69, Networks.py:12-32) plus the
 * constructor arguments of ULSTMnet2D (Networks.py:179) and the first-call shapes that freeze the stateful
 * ConvLSTM states (B,H,W). */
typedef struct lu_config {
  int32_t n_levels;
  int32_t n_lstm[LU_MAX_LEVELS];
  int32_t lstm_k[LU_MAX_LEVELS][LU_MAX_PER_LEVEL];
  int32_t lstm_f[LU_MAX_LEVELS][LU_MAX_PER_LEVEL];
  int32_t n_down[LU_MAX_LEVELS];
  int32_t down_k[LU_MAX_LEVELS][LU_MAX_PER_LEVEL];
  int32_t down_f[LU_MAX_LEVELS][LU_MAX_PER_LEVEL];
  int32_t n_up[LU_MAX_LEVELS];
  int32_t up_k[LU_MAX_LEVELS][LU_MAX_PER_LEVEL];
  int32_t up_f[LU_MAX_LEVELS][LU_MAX_PER_LEVEL];
  int32_t in_channels;     /* 1 for the CTC path */
  int32_t channels_first;  /* data_format[1]=='C' (Networks.py:181-182) */
  int32_t pad_image;       /* Networks.py:186,210 */
  int32_t batch, max_t, height, width;
  int32_t precision;       /* LU_PREC_* */
  int32_t engine;          /* LU_ENGINE_* */
  int32_t gate;            /* LU_GATE_* */
  int32_t a_mode;          /* LU_AMODE_* */
  int32_t train;           /* 1: allocate what forward(training=True)+backward need */
  float lrelu_alpha;       /* slope of the LeakyReLU after every BatchNorm; 0.3 = Keras-2 LeakyReLU() as the reference
                              constructs it (Networks.py:58,139) -- the caller must set it */
  /* stand-alone blocks (0 / LU_BLOCK_NET everywhere for the network).  LU_BLOCK_DOWN = DownBlock2D(conv_kernels,
   * lstm_kernels, stride, data_format) (Networks.py:37-75): n_levels = 1, level 0 of the lstm / down lists, in_channels,
   * height x width of the input, no padding.  LU_BLOCK_UP = UpBlock2D(kernels, up_factor, data_format, return_logits)
   * (Networks.py:124-153): n_levels = 1, entry 0 of the up list, in_channels / height / width of the LOW-resolution
   * input, skip_channels of the skip input (block_stride x larger), batch = frames, max_t = 1. */
  int32_t block_kind;      /* LU_BLOCK_* */
  int32_t block_stride;    /* DownBlock2D: stride of its first convolution (1 or 2); UpBlock2D: up_factor (1 or 2) */
  int32_t skip_channels;   /* UpBlock2D: channels of the skip input */
  int32_t return_logits;   /* UpBlock2D(return_logits=True): the last convolution's output, before BN (Networks.py:148-149) */
} lu_config;

typedef struct lu_handle_s* lu_handle;

const char* lu_last_error(void);
/* library/ABI version and whether this build is the CUDA product build (1) -- never 0 in a shipped .so */
int lu_version(void);
int lu_is_cuda_build(void);

/* ULSTMnet2D.__init__ (Networks.py:179-206): validates the level counts (ValueError -> non-zero), derives the
 * layer plan, parameter layout and workspace size.  No device memory is allocated. */
int lu_create(const lu_config* cfg, lu_handle* out);
int lu_destroy(lu_handle h);

/* device workspace (activations, recurrent states, packed bf16 weights, tables); caller allocates, library
 * zero-fills.  States live inside the workspace and persist across calls (stateful=True, Networks.py:48-50). */
int lu_workspace_bytes(lu_handle h, size_t* bytes);
int lu_bind_workspace(lu_handle h, void* dev_ws, size_t bytes, void* stream);

/* Keras variable layout (model.trainable_variables / save_weights, train2D.py:92,236): tensors in Keras
 * layouts (HWIO kernels, ConvLSTM (k,k,Cin,4F) gate order i,f,c,o) concatenated in one flat fp32 buffer,
 * trainable tensors first, BN moving statistics after them. */
int lu_param_count(lu_handle h, int32_t* n_tensors, int64_t* n_elements, int64_t* n_trainable_elements);
int lu_param_info(lu_handle h, int32_t idx, char* name, int32_t name_cap, int64_t* shape4, int32_t* rank,
                  int64_t* offset, int32_t* trainable);
int lu_bind_params(lu_handle h, float* dev_params);
/* call after the flat parameter buffer changed (load_weights / optimizer step): re-packs the tensor-core
 * operand copies of the weights */
int lu_params_changed(lu_handle h, void* stream);

/* ULSTMnet2D.call (Networks.py:208-254): x is (B,T,C,H,W) [channels_first] or (B,T,H,W,C); writes logits and
 * softmax of shape (B,T,3,H,W) / (B,T,H,W,3); mutates the recurrent states.  training selects BN batch
 * statistics (and keeps what backward needs when cfg.train). */
int lu_forward(lu_handle h, const float* dev_x, int32_t T, int32_t training, float* dev_logits,
               float* dev_softmax, void* stream);

/* DownBlock2D.call (Networks.py:60-75) / UpBlock2D.call (Networks.py:141-153) of a handle created with block_kind =
 * LU_BLOCK_DOWN / LU_BLOCK_UP.  DOWN: x is (B,T,C,H,W) / (B,T,H,W,C), dev_skip NULL; dev_out receives `activ`, the 4-D
 * (B*T, F, H/stride, W/stride) / (B*T, H/stride, W/stride, F) tensor the block returns second (its first return value is
 * the same data reshaped to 5-D); mutates the block's recurrent states.  UP: x is (N,C,h,w) / (N,h,w,C), skip
 * (N,Cs,f*h,f*w) / (N,f*h,f*w,Cs), T = 1; dev_out is (N,F,f*h,f*w) / (N,f*h,f*w,F).  lu_block_out_shape gives
 * {frames per time step (B or N), F, H_out, W_out}.  training selects BatchNorm batch statistics (forward only: block
 * handles have no backward). */
int lu_block_forward(lu_handle h, const float* dev_x, const float* dev_skip, int32_t T, int32_t training,
                     float* dev_out, void* stream);
int lu_block_out_shape(lu_handle h, int64_t* shape4);

/* Launch-bound shapes (Inference2D's per-frame call, B=1, T=1: Inference2D.py:59): replay the inference forward as an
 * instantiated CUDA graph (one per T and recurrent-state ping-pong parity; inputs / outputs pass through fixed staging
 * buffers of the workspace).  Available when batch*max_t <= 8 and the handle was not created for training;
 * *effective returns whether it is on.  Results are identical to the plain launches. */
int lu_set_graph_mode(lu_handle h, int32_t enable, int32_t* effective);

/* reset_states_per_batch (Networks.py:77-84,279-281): h,c *= mask[b]; mask is (B,) fp32 on the device */
int lu_reset_states(lu_handle h, const float* dev_mask, void* stream);
/* DownBlock2D.reset_states_per_batch (Networks.py:77-84) called on one block alone: the ConvLSTM layers of `level` only */
int lu_reset_level_states(lu_handle h, int32_t level, const float* dev_mask, void* stream);
/* get_states / set_states (Networks.py:86-98,283-291): one (B,F,H,W)/(B,H,W,F) fp32 tensor per call;
 * which: 0 = h, 1 = c.  dev_in == NULL zeroes the state (Keras reset_states(None)). */
int lu_state_shape(lu_handle h, int32_t level, int32_t layer, int64_t* shape4);
int lu_get_state(lu_handle h, int32_t level, int32_t layer, int32_t which, float* dev_out, void* stream);
int lu_set_state(lu_handle h, int32_t level, int32_t layer, int32_t which, const float* dev_in, void* stream);

/* WeightedCELoss (losses.py:13-27) + tape.gradient (train2D.py:89-92) for the last lu_forward(training=1):
 * labels (B,T,1,H,W)/(B,T,H,W,1) fp32 in {-1,0,1,2}; writes the scalar loss and the flat gradient of the
 * trainable prefix of the parameter buffer. */
int lu_loss_backward(lu_handle h, const float* dev_labels, const float* class_weights3, float* dev_loss,
                     float* dev_grads, void* stream);
/* Data-parallel training: `fn(offset, count, user)` is called on the host, from inside lu_loss_backward, as
 * soon as every launch that writes dev_grads[offset, offset+count) has been enqueued -- one call per Up / Down block,
 * decoder first.  The caller can start the all-reduce of that range on another stream (ordered after `stream` at that
 * point) while the rest of the backward runs.  NULL switches it off. */
typedef void (*lu_grad_bucket_fn)(int64_t offset, int64_t count, void* user);
int lu_set_grad_bucket_callback(lu_handle h, lu_grad_bucket_fn fn, void* user);
/* Synchronised BatchNorm for data-parallel training (SURVEY 8e option ii; the survey's `lu_bn_stats_{export,import}`):
 * the reference normalises with the statistics of the whole batch on one device; with the batch sharded over ranks
 * each BN layer's local moments are exported to `fn(dev_vec, count, user)` in the middle of lu_forward(training=1) /
 * lu_loss_backward -- the callee sums the fp64 vector over the ranks IN PLACE on the compute stream (one all-reduce of
 * 3 x channels doubles per layer forward, 2 x channels backward) -- and imported back as the global mean / variance
 * (forward) and the global means of g and g * xhat (backward).  Every rank must hold the same number of frames.
 * With the callback set the loss normaliser is global too (losses.py:26 divides by the valid pixels of the whole batch):
 * the valid-pixel count is summed over the ranks and loss / gradients are scaled by world_size, so that their MEAN over
 * the ranks equals the single-device value.  fn == NULL (default): local statistics, no collective in the forward. */
typedef void (*lu_bn_sync_fn)(double* dev_vec, int64_t count, void* user);
int lu_set_bn_sync_callback(lu_handle h, lu_bn_sync_fn fn, void* user, int32_t world_size);
/* optimizer.apply_gradients with Keras Adam (train2D.py:61,93): step is 1-based; m,v are flat fp32 buffers */
int lu_adam_step(lu_handle h, const float* dev_grads, float* dev_m, float* dev_v, float lr, float beta1,
                 float beta2, float eps, int64_t step, void* stream);

/* test hook: copy an internal NHWC buffer of the layer called `name` (e.g. "UpLayers/1/Conv/0") to fp32.
 * kind 0 = activation output, 1 = gradient of it (after lu_loss_backward), 2 = ConvLSTM gate pre-activation gradient */
int lu_debug_buffer(lu_handle h, const char* name, int32_t kind, float* dev_out, int64_t* shape4, void* stream);

/* counters for bench.py: kernels launched by this handle since the last reset */
int lu_launch_count(lu_handle h, int64_t* launches, int32_t reset);
/* algorithmic FLOPs (2*MAC, padding included) of one forward over T frames per sample at the bound shape */
int lu_forward_flops(lu_handle h, int32_t T, double* flops);
/* the ConvLSTM layers' share of lu_forward_flops (the dominant kernel's algorithmic work) */
int lu_lstm_flops(lu_handle h, int32_t T, double* flops);
/* timing of the dominant kernel for bench.py's roofline: returns the milliseconds and launch count accumulated by
 * CUDA-event pairs recorded around every ConvLSTM launch since the last call (synchronises on the last event),
 * then enables / disables the recording for the following forwards */
int lu_lstm_kernel_time(lu_handle h, int32_t enable, float* ms_total, int32_t* launches);
/* the same for every tensor-core kernel class (bench.py's `train` block names the dominant kernel of the training step):
 * ms4 / launches4 are indexed by LU_KC_*; lu_class_flops gives the algorithmic FLOPs (2*MAC) of each class for one
 * training step over T frames per sample at the bound batch size */
enum { LU_KC_LSTM_FWD = 0, LU_KC_CONV_FWD = 1, LU_KC_DGRAD = 2, LU_KC_WGRAD = 3, LU_KC_COUNT = 4 };
int lu_kernel_times(lu_handle h, int32_t enable, float* ms4, int32_t* launches4);
int lu_class_flops(lu_handle h, int32_t T, double* flops4);

/* ---- instance labelling of the soft-max maps (Inference2D.py:64-123): replaces the reference's numpy / SciPy / OpenCV
 * post-processing that follows the model call; results are bit-identical to it.  Stateless: everything
 * lives in the caller's workspace. */
typedef struct lu_post_params {
  float edge_thresh;        /* 0.2 in the reference (Inference2D.py:66) */
  int32_t edge_d2_limit;    /* an edge pixel joins the nearest cell if its SQUARED distance is < this; the host derives
                               it from params.edge_dist with the reference's float64 `sqrt(d2) < edge_dist` (:78) */
  int32_t min_cell_size;    /* params.min_cell_size / max_cell_size on the core area (:117-118) */
  int32_t max_cell_size;
  int32_t fov;              /* params.FOV (:94-104), 0 = off */
  int32_t channels_first;   /* 1: soft-max is (frames,3,H,W) [the reference's NCHW path]; 0: (frames,H,W,3) */
} lu_post_params;
int lu_post_workspace_bytes(int32_t frames, int32_t H, int32_t W, size_t* bytes);
/* dev_softmax: fp32 soft-max of `frames` frames; dev_labels: uint16 (frames,H,W) instance labels 1..kept, 0 = none;
 * dev_info (optional): int32 (frames,4) = {components found incl. background (cv2's count), labels kept,
 * 1 if the frame needed the sequential per-label pass, 0} */
int lu_postprocess(const float* dev_softmax, int32_t frames, int32_t H, int32_t W, const lu_post_params* params,
                   uint16_t* dev_labels, int32_t* dev_info, void* dev_ws, size_t ws_bytes, void* stream);
int lu_post_launch_count(int64_t* launches, int32_t reset);

/* ---- per-step metrics of the reference's train / validation step (train2D.py:97-102,111-116): the SEG measure
 * (losses.py:29-88: tf.py_function with SciPy labelling and Python loops on the host) and the sparse categorical
 * accuracy, computed on the device from the labels and logits that are already there.
 * dev_labels: (frames,1,H,W) == (frames,H,W,1) fp32 in {-1,0,1,2}; dev_logits: (frames,3,H,W) [channels_first] or
 * (frames,H,W,3).  dev_result4 (4 doubles): sum of the truth objects' scores, truth objects, correctly classified
 * pixels, pixels -- SEG = r[0]/r[1] (NaN without objects, like np.mean of nothing), accuracy = r[2]/r[3]. */
int lu_seg_workspace_bytes(int32_t frames, int32_t H, int32_t W, size_t* bytes);
int lu_seg_measure(const float* dev_labels, const float* dev_logits, int32_t frames, int32_t H, int32_t W,
                   int32_t channels_first, double* dev_result4, void* dev_ws, size_t ws_bytes, void* stream);

/* ---- training-reader augmentation on the device (CTCRAMReaderSequence2D._load_and_enqueue, DataHandeling.py:262-395,
 * and its static helpers :150-261): contrast / brightness, cv2.warpAffine + scipy map_coordinates elastic warp of image
 * and segmentation, _fix_transformed_segmentation, flips, rot90 -- one call per sequence chunk, results written where
 * the caller points (its slice of the (B,T,1,H,W) batch tensors).  Random numbers stay on the host (np.random, same
 * draws as the reference). */
typedef struct lu_aug_params {
  int32_t frames, H, W;
  int32_t randomize;        /* contrast / brightness on (self.randomize, :330-336) */
  int32_t elastic;          /* affine + elastic warp on (self.elastic_augmentation, :338-366) */
  int32_t flip0, flip1;     /* cv2.flip(., 0) / cv2.flip(., 1) (:369-374) */
  int32_t rot90;            /* np.rot90(., k) (:375-377); odd k needs H == W like the reference's fixed queue shapes */
  double affine[6];         /* the 2x3 matrix cv2.getAffineTransform returned (:150-168); inverted like cv2.warpAffine */
} lu_aug_params;
int lu_aug_workspace_bytes(int32_t frames, int32_t H, int32_t W, size_t* bytes);
/* dev_img / dev_seg: (frames,H,W) fp32 crops (seg: instance labels, -1 = not annotated); dev_contrast / dev_brightness:
 * (frames) fp32; dev_coords: (2,H,W) float64 from lu_elastic_coords (NULL without elastic); outputs (frames,H,W) fp32:
 * image, and segmentation in {-1, 0, 1, 2} */
int lu_augment_sequence(const float* dev_img, const float* dev_seg, const float* dev_contrast, const float* dev_brightness,
                        const double* dev_coords, const lu_aug_params* params, float* dev_img_out, float* dev_seg_out,
                        void* dev_ws, size_t ws_bytes, void* stream);
/* _get_indices4elastic_transform (:183-193): dev_rand (2,H,W) float64 uniform [0,1) (x field first), dev_weights the
 * 2*lw+1 normalised taps of scipy's gaussian_filter(sigma, truncate=4), dev_tmp (2,H,W) scratch -> dev_coords (2,H,W) =
 * (y + dy, x + dx) */
int lu_elastic_coords(const double* dev_rand, const double* dev_weights, int32_t lw, int32_t H, int32_t W, double alpha,
                      double* dev_tmp, double* dev_coords, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LSTM_UNET_B200_H */
