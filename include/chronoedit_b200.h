/*
 * chronoedit_b200 — C ABI of the B200-native ChronoEdit denoising hot path.
 *
 * The reference (nv-tlabs/ChronoEdit) has no FFI for this path: the boundary is two Python objects registered
 * into a diffusers pipeline (pipeline_chronoedit.py:175-183).  This header is the C-ABI underneath the Python
 * mirror classes in chronoedit_b200/ (INTEGRATION.md shows the ctypes binding); each entry point names the
 * reference interface it stands in for.
 *
 * Conventions
 *   - every pointer argument is a DEVICE pointer owned by the caller unless the name ends in _host;
 *   - activations / weights are bf16 unless stated; tensors are dense row-major in the layouts given;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it and never synchronise the device
 *     (the *_host variants synchronise `stream` before returning because they hand results back in host memory);
 *   - return value 0 = success, negative = failure (ce_last_error() holds the message of the calling thread's last
 *     failure); there is NO CPU fallback: on a machine without an sm_100 GPU every compute entry point fails;
 *   - one handle may be used from one thread at a time; distinct handles are independent.
 */
#ifndef CHRONOEDIT_B200_H_
#define CHRONOEDIT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CE_ABI_VERSION 1

int ce_abi_version(void);
const char* ce_last_error(void);
/* 0 iff the current CUDA device is sm_100 (B200). */
int ce_device_check(void);

/* ------------------------------------------------------------------------------------------------------------
 * DiT: ChronoEditTransformer3DModel  (chronoedit_diffusers/transformer_chronoedit.py:298-476)
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct ce_dit ce_dit;

typedef struct ce_dit_config {   /* = the @register_to_config arguments, transformer_chronoedit.py:342-360 */
  int32_t num_attention_heads;   /* 40 */
  int32_t attention_head_dim;    /* 128 (only value built) */
  int32_t in_channels;           /* 36 */
  int32_t out_channels;          /* 16 */
  int32_t text_dim;              /* 4096 */
  int32_t freq_dim;              /* 256 */
  int32_t ffn_dim;               /* 13824 */
  int32_t num_layers;            /* 40 */
  int32_t image_dim;             /* 1280; 0 = no image embedder / no added k,v projections */
  int32_t added_kv_proj_dim;     /* 5120 */
  int32_t rope_max_seq_len;      /* 1024 */
  int32_t rope_temporal_skip_len;/* 8 */
  float eps;                     /* 1e-6 */
  int32_t patch_t, patch_h, patch_w; /* (1,2,2) (only value built) */
} ce_dit_config;

/* ChronoEditTransformer3DModel.__init__ (:342-395).  No weights are allocated: the caller owns them. */
int ce_dit_create(const ce_dit_config* cfg, ce_dit** out);
void ce_dit_destroy(ce_dit* h);

/* Register one parameter by name (device pointer, kept alive by the caller until ce_dit_destroy).
 * Names are the reference state_dict's (e.g. "blocks.3.ffn.net.2.weight") except for the fused / stacked buffers the
 * Python mirror builds as views over the reference parameters:
 *   blocks.N.attn1.to_qkv.{weight,bias}      [3D,D] / [3D]   rows = to_q | to_k | to_v
 *   blocks.N.attn2.to_kv.{weight,bias}       [2D,D] / [2D]   rows = to_k | to_v
 *   blocks.N.attn2.add_kv_proj.{weight,bias} [2D,Da]/ [2D]   rows = add_k_proj | add_v_proj
 *   blocks.scale_shift_table                 [num_layers,6,D] fp32 (stack of blocks.N.scale_shift_table)
 *   patch_embedding.weight                   [D, Cin*pt*ph*pw] (the Conv3d weight, flattened)
 * dtype: 0 = bf16, 1 = fp32.  fp32 is required for (and only for) the reference's `_keep_in_fp32_modules`
 * (time_embedder, scale_shift_table, norm1/norm2/norm3 affine parameters, :338). */
int ce_dit_set_weight(ce_dit* h, const char* name, const void* ptr, int dtype, int64_t numel);
/* 0 when every parameter the configuration needs has been registered; otherwise fails and names the first missing. */
int ce_dit_weights_complete(const ce_dit* h);

/* Scratch bytes one forward needs for `batch` samples of latent geometry (frames, height, width) and text_len tokens. */
int64_t ce_dit_workspace_bytes(const ce_dit* h, int batch, int frames, int height, int width, int text_len);

/* ChronoEditTransformer3DModel.forward (:397-476).
 *   hidden_states [B, in_channels, frames, height, width]   timestep [B] fp32 (the pipeline's int64 t cast to float, :155-157)
 *   encoder_hidden_states [B, text_len, text_dim]           encoder_hidden_states_image [B, 257, image_dim] or NULL
 *   sample (out) [B, out_channels, frames, height, width]
 * frames must be 2 or rope_temporal_skip_len (the reference asserts this, :205).
 * intermediates: optional; when non-NULL receives block 0's output [B*L, D] bf16 (parity tests). */
int ce_dit_forward(ce_dit* h, const void* hidden_states, const float* timestep, const void* encoder_hidden_states,
                   const void* encoder_hidden_states_image, void* sample, int batch, int frames, int height, int width,
                   int text_len, void* workspace, int64_t workspace_bytes, void* block0_out, void* stream);

/* The same forward with the STEP-INVARIANT context kept across calls: the text / image embedders (:147-165) and every block's
 * cross-attention K/V projections + RMSNorm (:52-60, :84-95) depend only on encoder_hidden_states[_image] and the weights, not
 * on the latents or the timestep, so within one edit they are the same at every denoising step.
 *   ctx_cache: device buffer of >= ce_dit_context_cache_bytes(h, batch, text_len) bytes owned by the caller, or NULL (= ce_dit_forward);
 *   ctx_reuse = 0: compute the context and leave it in ctx_cache;  ctx_reuse = 1: the cache already holds the context of THESE
 *   encoder states and weights (the caller's promise) -- the embedders and the 2 x num_layers K/V GEMMs are skipped.
 * The result is bit-identical either way (same kernels, same inputs).  Algorithmic FLOP counts in bench.py stay un-hoisted. */
int64_t ce_dit_context_cache_bytes(const ce_dit* h, int batch, int text_len);
int ce_dit_forward_ex(ce_dit* h, const void* hidden_states, const float* timestep, const void* encoder_hidden_states,
                      const void* encoder_hidden_states_image, void* sample, int batch, int frames, int height, int width,
                      int text_len, void* workspace, int64_t workspace_bytes, void* block0_out, void* ctx_cache,
                      int64_t ctx_cache_bytes, int ctx_reuse, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Sequence parallelism over the GPUs of one node for single-edit latency (reference: the Ulysses /
 * xfuser path chronoedit_diffsynth/wan_video_new_chronoedit.py:330-355, 1448-1498 and the TE ring of
 * chronoedit/_src/networks/wan2pt1.py:352-353, 917-941).  One process per GPU, all weights on every rank, the tokens of the edit
 * split into `world` contiguous ranges.  The heads <-> tokens exchange around the self-attention is done by the producing kernels
 * themselves with stores into the consumer's memory over NVLink (csrc/seqpar.cuh); phases are ordered by a flag barrier in peer
 * memory.  Every rank passes the SAME full inputs and receives the same full sample, bit-identical to the single-GPU forward.
 *   region: one device buffer of ce_dit_sp_region_bytes(...) per rank, allocated with ce_ipc_alloc (zeroed), shared with the peers
 *   through ce_ipc_get_handle / ce_ipc_open (CUDA IPC; the 64-byte handles travel by any host channel, e.g. torch.distributed);
 *   region_ptrs[w] = address of rank w's region in THIS process.  world = 1 switches it off.
 * ------------------------------------------------------------------------------------------------------------ */
int64_t ce_dit_sp_region_bytes(const ce_dit* h, int batch, int frames, int height, int width, int world);
int ce_dit_sp_configure(ce_dit* h, int rank, int world, void* const* region_ptrs, int64_t region_bytes);
int ce_ipc_alloc(int64_t bytes, void** ptr);
int ce_ipc_free(void* ptr);
int ce_ipc_get_handle(void* ptr, void* handle_out_64_bytes);
int ce_ipc_open(const void* handle_64_bytes, void** ptr);
int ce_ipc_close(void* ptr);

/* Parity aid: from now on every forward copies the output of block layers[i] ([B*L, D] bf16) to dst[i] (n = 0 clears). */
int ce_dit_set_capture(ce_dit* h, const int32_t* layers, void* const* dst, int n);

/* Same call with every tensor in (pinned) HOST memory: H2D copies of the inputs, forward, D2H of the sample, all on
 * `stream`, which is synchronised before returning.  `staging` is a device buffer of at least
 * ce_dit_host_staging_bytes(...) bytes.  This is the end-to-end ("e2e") path bench.py times. */
int64_t ce_dit_host_staging_bytes(const ce_dit* h, int batch, int frames, int height, int width, int text_len);
int ce_dit_forward_host(ce_dit* h, const void* hidden_states_host, const float* timestep_host,
                        const void* encoder_hidden_states_host, const void* encoder_hidden_states_image_host,
                        void* sample_host, int batch, int frames, int height, int width, int text_len, void* staging,
                        int64_t staging_bytes, void* workspace, int64_t workspace_bytes, void* stream);

/* The host-buffer call with the step-invariant context cache of ce_dit_forward_ex (device buffer, caller-owned).  The sample also
 * stays in `staging` at byte offset ce_dit_host_staging_sample_offset(...) (bf16 [B, out_channels, frames, height, width]) until
 * the next call, so that the scheduler step can consume it on the device without a second copy. */
int ce_dit_forward_host_ex(ce_dit* h, const void* hidden_states_host, const float* timestep_host,
                           const void* encoder_hidden_states_host, const void* encoder_hidden_states_image_host,
                           void* sample_host, int batch, int frames, int height, int width, int text_len, void* staging,
                           int64_t staging_bytes, void* workspace, int64_t workspace_bytes, void* ctx_cache,
                           int64_t ctx_cache_bytes, int ctx_reuse, void* stream);
int64_t ce_dit_host_staging_sample_offset(const ce_dit* h, int batch, int frames, int height, int width, int text_len);

/* VALIDATION mode: the same forward in fp32 -- fp32 inputs, fp32 parameters (registered under their REFERENCE names, unfused:
 * blocks.N.attn1.to_q.weight ..., blocks.N.scale_shift_table [6*D], scale_shift_table [2*D], patch_embedding.weight [D, Cin*4]),
 * fp32 residual stream, fp32 outputs.  Matrix products run on the same tcgen05 GEMM with every fp32 operand split into two bf16
 * terms (A_hi.W_hi + A_lo.W_hi + A_hi.W_lo, fp32 accumulate; csrc/dit_fp32.cu).  It exists so that north_star's
 * rtol 1e-3 / atol 1e-4 can be asserted end to end against the reference's fp32 run (tests/test_gpu_fp32_mode.py); it is about
 * 5-10x slower than the bf16 path and is not what bench.py measures.  block_out: optional, block 0's output [B*L, D] fp32. */
int ce_dit_set_weight_fp32(ce_dit* h, const char* name, const float* ptr, int64_t numel);
int64_t ce_dit_fp32_workspace_bytes(const ce_dit* h, int batch, int frames, int height, int width, int text_len);
int ce_dit_forward_fp32(ce_dit* h, const float* hidden_states, const float* timestep, const float* encoder_hidden_states,
                        const float* encoder_hidden_states_image, float* sample, int batch, int frames, int height, int width,
                        int text_len, void* workspace, int64_t workspace_bytes, float* block_out, void* stream);

/* Kernel launches issued by the last ce_dit_forward on this handle (bench.py's "gpu_launches"). */
int64_t ce_dit_last_launch_count(const ce_dit* h);

/* Device-side timing of the kernels of subsequent forwards: a CUDA event pair on the launch stream around every
 * launch (up to max_launches; later launches are not recorded).  ce_dit_profile_end waits for the last recorded event
 * and returns, per category c in {0: tcgen05 GEMM, 1: tcgen05 attention, 2: LayerNorm / RMSNorm row kernels,
 * 3: everything else}: total milliseconds, total algorithmic work (FLOPs for 0/1, bytes for 2) and launch count
 * (each array has 4 entries).  This is what bench.py's "roofline" object is computed from. */
int ce_dit_profile_begin(ce_dit* h, int max_launches);
int ce_dit_profile_end(ce_dit* h, double* ms_out, double* work_out, int64_t* count_out);

/* ------------------------------------------------------------------------------------------------------------
 * VAE: AutoencoderKLWan encode / decode as the pipeline uses them (pipeline_chronoedit.py:436-443, 776-781);
 * arithmetic = WanVAE_ of chronoedit/_src/tokenizers/wan2pt1.py:467-581 WITHOUT the latent mean/std (the pipeline applies
 * those itself, :427-445, :765-774).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct ce_vae ce_vae;

typedef struct ce_vae_config {      /* _video_vae cfg, wan2pt1.py:597-605 */
  int32_t dim;                      /* 96 */
  int32_t z_dim;                    /* 16 */
  int32_t dim_mult[4];              /* 1,2,4,4 */
  int32_t num_res_blocks;           /* 2 */
  int32_t temporal_downsample[3];   /* 0,1,1 */
} ce_vae_config;

int ce_vae_create(const ce_vae_config* cfg, ce_vae** out);
void ce_vae_destroy(ce_vae* h);
/* Parameters by the in-tree twin's names (e.g. "decoder.upsamples.3.time_conv.weight"), all bf16, PRE-PACKED by the caller:
 *   convolution weights [Cout, Cin, kt, kh, kw] -> [Cout, taps, Cin_pad] with taps ordered (dt, dh, dw), Cin_pad = Cin rounded
 *     up to 64 (zero filled) when Cin >= 64; for Cin < 64 -> [Cout, Kpad] with K = taps*Cin rounded up to 8;
 *   "*.gamma" [C]; attention "to_qkv.weight" [3C, C], "proj.weight" [C, C]; biases [Cout]. */
int ce_vae_set_weight(ce_vae* h, const char* name, const void* ptr, int64_t numel);
/* decode != 0: (frames, height, width) is the LATENT geometry; else the PIXEL geometry. */
int64_t ce_vae_workspace_bytes(ce_vae* h, int decode, int frames, int height, int width);
/* video [3, frames, height, width] (planar, one sample, values in [-1,1], frames = 1+4k) -> posterior moments
 * [2*z_dim, 1+k, height/8, width/8]: channels [0,z) = mean (= AutoencoderKLWan.encode(x).latent_dist.mode()), [z,2z) = logvar */
int ce_vae_encode(ce_vae* h, const void* video, void* moments, int frames, int height, int width, void* workspace,
                  int64_t workspace_bytes, void* stream);
/* z [z_dim, Tl, h, w] -> video [3, 1+4(Tl-1), 8h, 8w]; clamp != 0 clamps to [-1,1] (= AutoencoderKLWan.decode(z)[0]) */
int ce_vae_decode(ce_vae* h, const void* z, void* video, int latent_frames, int latent_height, int latent_width, int clamp,
                  void* workspace, int64_t workspace_bytes, void* stream);
int64_t ce_vae_last_launch_count(const ce_vae* h);

/* ------------------------------------------------------------------------------------------------------------
 * Encoders in front of the loop, once per edit:
 *   transformers.UMT5EncoderModel   -- pipeline_chronoedit.py:205-244  self.text_encoder(ids, mask).last_hidden_state
 *   transformers.CLIPVisionModel    -- pipeline_chronoedit.py:246-254  self.image_encoder(**image, output_hidden_states=True).hidden_states[-2]
 * Parameters are registered by their transformers state_dict names, bf16, with these host-side fusions:
 *   UMT5: "encoder.block.N.layer.0.SelfAttention.qk.weight" [2*H*d_kv, d_model] = rows of q | k;   "shared.weight" [vocab, d_model]
 *   CLIP: "vision_model.encoder.layers.N.self_attn.qk_proj.{weight,bias}" = q | k;  patch_embedding.weight flattened [D, 3*ps*ps] with K padded
 *         to a multiple of 8;  every LayerNorm additionally as fp32 "<name>.weight_f32" / "<name>.bias_f32" [D].
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct ce_encoder ce_encoder;
typedef struct ce_encoder_config {
  int32_t kind;          /* 0 = UMT5 text encoder, 1 = CLIP vision encoder */
  int32_t vocab_size;    /* UMT5: 256384 */
  int32_t d_model;       /* 4096 | 1280 */
  int32_t d_kv;          /* head dim: 64 | 80 */
  int32_t d_ff;          /* 10240 | 5120 */
  int32_t num_layers;    /* 24 | 32 */
  int32_t num_heads;     /* 64 | 16 */
  float eps;             /* 1e-6 | 1e-5 */
  int32_t image_size;    /* CLIP: 224 */
  int32_t patch_size;    /* CLIP: 14 */
  int32_t hidden_act;    /* CLIP: 0 = quick_gelu, 1 = gelu (erf) */
} ce_encoder_config;
int ce_encoder_create(const ce_encoder_config* cfg, ce_encoder** out);
void ce_encoder_destroy(ce_encoder* h);
int ce_encoder_set_weight(ce_encoder* h, const char* name, const void* ptr, int64_t numel);
int64_t ce_encoder_workspace_bytes(const ce_encoder* h, int batch, int seq_len);   /* CLIP: seq_len = 1 + (image_size/patch_size)^2 */
int64_t ce_encoder_last_launch_count(const ce_encoder* h);
/* input_ids [B, L] int64 (device); valid_len_host[b] = number of leading unmasked tokens (HOST array; the tokenizer pads on the right);
 * bias_tables [num_layers, H, 2L-1] bf16 (device): relative_attention_bias of layer l, head h at relative position (key - query) + L - 1
 * (the bucket function is host code, chronoedit_b200/encoders.py);  last_hidden_state [B, L, d_model] bf16.  L must be a multiple of 8. */
int ce_umt5_encode(ce_encoder* h, const int64_t* input_ids, const int32_t* valid_len_host, void* last_hidden_state, int batch,
                   int seq_len, const void* bias_tables, void* workspace, int64_t workspace_bytes, void* stream);
/* pixel_values [B, 3, S, S] bf16 -> hidden state after `layers_to_run` encoder layers [B, 1 + (S/ps)^2, d_model] bf16
 * (hidden_states[-2] = num_layers - 1 layers; 0 = the embeddings after pre_layrnorm). */
int ce_clip_vision_encode(ce_encoder* h, const void* pixel_values, void* out, int batch, int layers_to_run, void* workspace,
                          int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Individual hot-path operators (used by the handles above; exported for operator-level parity tests / profiling)
 * ------------------------------------------------------------------------------------------------------------ */

/* One channels-last convolution on the tcgen05 implicit-GEMM kernel (CausalConv3d / Conv2d of wan2pt1.py:42-60, 98-110):
 *   y[to,oh,ow,n] = bias[n] + sum W[n,(dt,dh,dw),c] * x[t_base + to*st + dt, oh*sh + dh - ph, ow*sw + dw - pw, c]
 * x [Tin,Hin,Win,Cin], w packed [Cout, taps, Cin_pad64], y [Tout,Hout,Wout,Cout] (split_time: [2*Tout,...,Cout/2]),
 * resid optional (same geometry as y). */
int ce_conv3d_cl_bf16(const void* x, int Tin, int Hin, int Win, int Cin, const void* w, const void* bias, int Cout, int kt,
                      int kh, int kw, int st, int sh, int sw, int ph, int pw, int t_base, void* y, int Tout, int Hout, int Wout,
                      const void* resid, int split_time, void* stream);

/* out[M,N] = epilogue(A[M,K] x W[N,K]^T + bias): one nn.Linear (+ fused elementwise tail).
 * epilogue: 0 bias | 1 bias+GELU(tanh) | 2 bias+GELU(erf) | 3 bias, gate*y + resid (fp32) | 4 bias, y + resid
 * out_f32 (optional, [M,N]): acc + bias in fp32 before any rounding. */
int ce_linear_bf16(const void* A, int lda, const void* W, int ldw, const void* bias, void* out, int ldo, float* out_f32,
                   int M, int N, int K, int epilogue, const void* resid, int ldr, const float* gate, int gate_stride,
                   int rows_per_batch, void* stream);

/* F.scaled_dot_product_attention (non-causal, no mask, head_dim 128) on [B, L, H*128]-strided q/k/v
 * (transformer_chronoedit.py:91-99).  accumulate=1 adds into `out` in bf16 (text + image cross-attention, :103-104). */
int ce_attention_bf16(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int B,
                      int H, int Lq, int Lk, float scale, int accumulate, void* stream);

/* Profiling aid: subsequent ce_attention_bf16 calls record per-phase cycle counts of one softmax warp of block 0 into `buf`
 * (device, 16 x int64); NULL stops. */
int ce_debug_attention_timing(long long* buf);

/* Test / A-B aid: select the kernel that serves long single-source (self-)attention from now on: 6 = two query tiles per CTA, two
 * softmax threads per score row (default), 2 = two query tiles per CTA, one thread per row, 5 = cta_group::2 cluster kernel
 * (experimental), 0 = single-tile kernel, -1 = back to the default / CE_ATTN_V2. */
int ce_debug_attention_kernel(int version);

/* Cross-attention with TWO key/value sources in one launch: out = bf16(SDPA(q,k,v)) + bf16(SDPA(q,k2,v2)) — the text and
 * image streams of ChronoEditAttnProcessor2_0 (transformer_chronoedit.py:84-104). */
int ce_attention_dual_bf16(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const void* k2, int ldk2,
                           const void* v2, int ldv2, void* out, int ldo, int B, int H, int Lq, int Lk, int Lk2, float scale,
                           void* stream);

/* FP32LayerNorm (+ adaLN modulate or affine) over the last dim of x[rows, D] -> y (bf16). (:279, :284, :289, :460) */
int ce_layernorm_bf16(const void* x, int ldx, void* y, int ldy, int rows, int D, float eps, const float* scale,
                      const float* shift, int mod_stride, int rows_per_batch, const float* weight, const float* bias,
                      void* stream);

/* diffusers RMSNorm across heads, then optional interleaved RoPE, in place (:62-65, :71-79).
 * rope_cos / rope_sin: fp32 [L, head_dim/2] or NULL. */
int ce_rmsnorm_rope_bf16(void* x, int ldx, int rows, int D, float eps, const void* weight, const float* rope_cos,
                         const float* rope_sin, int L, int head_dim, void* stream);

/* ChronoEditRotaryPosEmbed.forward (:168-213) as fp32 cos/sin tables [L, head_dim/2] written to HOST arrays. */
int ce_rope_table_host(int head_dim, int frames, int height_patches, int width_patches, int max_seq_len,
                       int temporal_skip_len, float theta, float* cos_out_host, float* sin_out_host);

/* ------------------------------------------------------------------------------------------------------------
 * Sampling glue either side of the DiT call: one launch per denoising step for
 *   noise = uncond + g*(cond - uncond)                                   chronoedit_diffusers/pipeline_chronoedit.py:736
 *   FlowUniPCMultistepScheduler.step: x0 conversion, UniC corrector, UniP predictor (solver_order 2, bh2, predict_x0,
 *   flow_prediction, lower_order_final)                                  chronoedit/_src/models/fm_solvers_unipc.py:670-756
 *   channels [0, c_lat) of the next cat([latents, condition], 1).to(bf16) pipeline_chronoedit.py:712
 * Every intermediate is rounded to the tensor dtype exactly where the reference's separate torch kernels round, so the
 * result is bit-identical to running the reference on CUDA.  The scalar coefficients are the caller's business (the Python
 * mirror chronoedit_b200/scheduler.py computes them with the same fp32 CPU tensor ops as the reference, :418-447, :565-620);
 * divisions by r_k arrive as the fp32 reciprocal because that is how torch's CUDA `tensor / scalar` evaluates.
 * ------------------------------------------------------------------------------------------------------------ */
#define CE_DTYPE_F32 0
#define CE_DTYPE_BF16 1

typedef struct ce_unipc_step_args {
  int32_t sample_dtype;          /* dtype of sample / last_sample / m_prev* / outputs: CE_DTYPE_F32 | CE_DTYPE_BF16 */
  int32_t model_dtype;           /* dtype of cond / uncond; (f32,f32) (f32,bf16) (bf16,bf16) are built */
  int64_t n;                     /* elements of the latent [B, c_lat, T, H, W] */
  const void* cond;              /* model output (conditional) */
  const void* uncond;            /* model output (unconditional) or NULL: no guidance */
  float guidance;
  float sigma;                   /* sigmas[step_index]: x0 = sample - sigma * v                       (:340-341) */
  const void* sample;            /* x_t */
  const void* last_sample;       /* sample before the previous predictor; corrector only                (:567) */
  const void* m_prev;            /* model_outputs[-1] before this call (x0 prediction of the previous step) or NULL */
  const void* m_prev2;           /* model_outputs[-2] before this call or NULL */
  int32_t use_corrector;         /* step_index > 0 and last_sample is set                               (:701-705) */
  int32_t c_order;               /* 1 | 2: this_order of the previous step */
  float c_x, c_m0, c_bh;         /* sigma_t/sigma_s0, alpha_t*h_phi_1, alpha_t*B_h at (step_index, step_index-1) */
  float c_inv_rk, c_rho0, c_rho1;/* order 2: 1/r_k and solve(R, b) cast to the sample dtype; order 1 uses 0.5 */
  int32_t p_order;               /* 1 | 2: min(2, steps - step_index, lower_order_nums + 1)             (:729-737) */
  float p_x, p_m0, p_bh;         /* the same three scalars at (step_index+1, step_index) */
  float p_inv_rk;                /* order 2 */
  float p_zero;                  /* order 1: alpha_t*B_h*0, still subtracted by the reference (sign of zero) */
  void* x0_out;                  /* -> model_outputs[-1] */
  void* corrected_out;           /* -> last_sample (written only when use_corrector) */
  void* prev_sample_out;         /* -> returned sample */
  void* model_input_out;         /* optional bf16 [B, c_total, inner]: channels [0, c_lat) <- prev_sample; NULL = skip */
  int64_t inner;                 /* T*H*W */
  int32_t c_lat, c_total;        /* 16, 36 */
} ce_unipc_step_args;

int ce_unipc_step(const ce_unipc_step_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CHRONOEDIT_B200_H_ */
