/* sdd_b200.h -- C ABI of the B200-native SuperDiff sampling hot path.
 *
 * The reference (mo-rsa24/super-diff-disease) has NO plugin / FFI boundary for this path:
 * it is pure Python over stock PyTorch.  The entry points below are
 * therefore what a binding for the reference's Python objects needs, one per reference call:
 *
 *   sdd_unet_create / sdd_unet_forward   <->  UNet.__init__ / UNet.forward     src/models/unet.py:38,57
 *   sdd_sampler_* (M = 1)                <->  DDPM.sample                      src/models/ddpm.py:31-45
 *   sdd_sampler_* (M >= 2)               <->  src/sampling.py (0 bytes in the reference; the
 *                                             superposed sampler the README describes, README.md:5,9)
 *   sdd_superpose_update                 <->  the per-step x update, ddpm.py:42-44, extended with
 *                                             the kappa softmax / Ito log-density increment (A7)
 *
 * Conventions: plain pointers and sizes only.  All data pointers are DEVICE pointers unless the
 * parameter name ends in _host.  `stream` is a cudaStream_t passed as void*.  Every function
 * returns 0 on success or an SDD_E* code; sdd_last_error() gives the message (thread-local).
 * Nothing here falls back to the CPU: without an sm_100 device every call fails with SDD_ENODEV.
 * The library never frees or retains caller memory beyond the call, except where stated.
 */
#ifndef SDD_B200_H
#define SDD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDD_ABI_VERSION 2

enum {
  SDD_OK = 0,
  SDD_EINVAL = 1,   /* bad argument / unsupported shape */
  SDD_ENODEV = 2,   /* no CUDA device of compute capability 10.x */
  SDD_ECUDA = 3,    /* a CUDA runtime / driver call failed */
  SDD_ENOMEM = 4
};

/* Number of fp32 tensors in a reference UNet state_dict (unet.py:40-55) and their order:
 * exactly the order torch's state_dict() yields for the reference module:
 *   time_mlp.1.{weight,bias}, time_mlp.3.{weight,bias}, then for each block in
 *   downs.0, downs.1, mid, ups.0, ups.1:
 *     block.0.{weight,bias} (GN1), block.2.{weight,bias} (conv1, [Cout,Cin,3,3]),
 *     block.3.{weight,bias} (GN2), block.5.{weight,bias} (conv2), time_emb.{weight,bias}. */
#define SDD_UNET_NUM_TENSORS 54

typedef struct sdd_unet sdd_unet_t;
typedef struct sdd_sampler sdd_sampler_t;

int sdd_abi_version(void);
const char* sdd_last_error(void);
/* 0 iff the current CUDA device is compute capability 10.x. */
int sdd_device_check(void);

/* ---- UNet (unet.py:37-65; default architecture only: in=out=1, time_emb_dim=256, base=64) ---- */

/* Build device-resident kernel-layout weights (fp16 conv operands, fp32 everything else) from the
 * 54 fp32 state-dict tensors (device pointers, contiguous, reference shapes).  The inputs are only
 * read during the call (stream-ordered); the handle owns its own copies. */
int sdd_unet_create(sdd_unet_t** out, const float* const* tensors, int num_tensors, void* stream);
int sdd_unet_destroy(sdd_unet_t* u);
/* Cap the number of samples one pass of the forward processes at a time (0 = automatic: up to 1.5 GB per ping-pong
 * activation buffer).  Chunking changes no result bit; it bounds the workspace.  Takes effect at the next call. */
int sdd_unet_set_max_chunk(sdd_unet_t* u, int max_samples);

/* eps_out[B,1,H,W] = UNet(x[B,1,H,W], t[B]); fp32 in/out, t is int64 as in unet.py:57.
 * H % 16 == 0 and W % 8 == 0 are required (SDD_EINVAL otherwise).  Workspace is owned by the
 * handle and grown on demand (so the first call at a new shape is not graph-capturable). */
int sdd_unet_forward(sdd_unet_t* u, const float* x, const int64_t* t, float* eps_out,
                     int B, int H, int W, void* stream);
/* Same, with the first layer's GroupNorm(1,1) statistics of x supplied by the caller: xstats[B,2] = (mean, rstd) per
 * sample, e.g. the xstats_out of the sdd_superpose_update call that produced x (NULL: computed here).  A step-by-step
 * loop driven through the operator entry points then reproduces sdd_sampler_run bit for bit. */
int sdd_unet_forward_xstats(sdd_unet_t* u, const float* x, const float* xstats, const int64_t* t, float* eps_out,
                            int B, int H, int W, void* stream);

/* ---- Extension (NO reference code -- the reference UNet, unet.py:37-65, is five full-resolution
 * blocks with no attention, resampling, skips or class input; oracle = oracle/unet_attn_oracle.py):
 * class-conditional multi-resolution UNet with self-attention at R/8 and R/16 (32^2 / 16^2 at R = 256), built from the
 * reference's own ResidualBlock (unet.py:18-34) so that every conv runs on the kernels of the reference path:
 *   enc0 RB(1,64)@R | pool | enc1 RB(64,128)@R/2 | pool | enc2 RB(128,128)@R/4 | pool | enc3 RB(128,128)+Attn@R/8 | pool |
 *   enc4 RB(128,128)+Attn@R/16 | mid RB(128,128)+Attn@R/16 | up+skip(enc3) | dec0 RB(128,128)+Attn@R/8 | up+skip(enc2) |
 *   dec1 RB(128,128)@R/4 | up+skip(enc1) | dec2 RB(128,64)@R/2 | up+skip(enc0) | out RB(64,1)@R
 * pool = 2x2 average, up = nearest neighbour, skips are ADDED; Attn = x + proj(softmax(q k^T / 8) v), [q|k|v] =
 * GroupNorm(4,128)(x) W_qkv^T + b, 2 heads of 64; the time embedding of every block is time_mlp(t) + class_emb[y].
 * tensors (129, fp32, device): time_mlp.1.{weight,bias}, time_mlp.3.{weight,bias}, class_emb.weight [num_classes,256];
 * the ten blocks in the order above, each as in the reference (block.0, block.2, block.3, block.5, time_emb: 10 tensors);
 * the four attention blocks (after enc3, enc4, mid, dec0), each norm.{weight,bias}, qkv.{weight [384,128],bias},
 * proj.{weight [128,128],bias}.  H % 256 == 0 and W % 128 == 0.  The handle is a sdd_unet_t: sdd_unet_forward* /
 * sdd_sampler_* accept it like a reference UNet (sdd_sampler_create conditions on the handle's label). */
#define SDD_UNET_ATTN_NUM_TENSORS 129
int sdd_unet_attn_create(sdd_unet_t** out, const float* const* tensors, int num_tensors, int num_classes, void* stream);
/* The class a sampler built on this handle (and a forward without per-sample labels) conditions on; default 0. */
int sdd_unet_set_label(sdd_unet_t* u, int label);
/* sdd_unet_forward_xstats with per-sample class labels y[B] (int64; NULL: the handle's label).  Variant handles only. */
int sdd_unet_forward_labeled(sdd_unet_t* u, const float* x, const float* xstats, const int64_t* t, const int64_t* y,
                             float* eps_out, int B, int H, int W, void* stream);

/* ---- Fused superposition update (A7): ONE kernel launch and one HBM pass per step ----
 * kappa = softmax_m(temperature * logq[b,:] + bias);  eps_bar = sum_m kappa_m eps[m,b,:]
 * x_out = alpha^-1/2 (x_in - (1-alpha)/sqrt(1-alpha_bar) eps_bar) + sqrt(beta) z        (ddpm.py:42-44)
 * logq[b,m] += <s_m, x_out-x_in> - beta D/2 - beta/2 <x_in, s_m> - beta/2 |s_m|^2,  s_m = -eps_m/sqrt(1-alpha_bar)
 * z: noise[B,D] if non-NULL; else Philox4x32-10 keyed (seed; element/4, sample_offset+b, draw_index)
 * if draw_index >= 0; else zero (the t == 0 step, ddpm.py:36).
 * x_out may alias x_in.  kappa_out[B,M] / logq_out[B,M] / xstats_out[B,2] (mean, rstd of x_out for the
 * next GroupNorm(1,1)) may be NULL; logq_out may alias logq.  M <= 4.  D % 4 == 0.
 * workspace: sdd_superpose_update_workspace(B, D, M) bytes, zero-initialised once by the caller (it holds the
 * per-segment partial sums and the arrival counters that elect each sample's finalising CTA; the kernel leaves the
 * counters at zero again). */
size_t sdd_superpose_update_workspace(int B, int D, int M);
int sdd_superpose_update(const float* x_in, float* x_out, const float* eps, const float* noise,
                         const float* logq, float* logq_out, float* kappa_out, float* xstats_out,
                         int B, int D, int M, float alpha, float alpha_bar, float beta,
                         float temperature, const float* bias,
                         uint64_t seed, int64_t sample_offset, int draw_index,
                         void* workspace, size_t workspace_bytes, void* stream);

/* SuperDiff "AND" step: same contract as sdd_superpose_update, but kappa[b,:] solves
 *   sum_j kappa_j = 1,   inc_i(kappa) = inc_0(kappa)  for i = 1..M-1
 * where inc_i is model i's log q increment of THIS step (affine in kappa through x_out - x_in); a singular system
 * (identical models) gives kappa = 1/M.  Two passes over x / eps / z (Gram reductions, then the update). */
int sdd_superpose_update_and(const float* x_in, float* x_out, const float* eps, const float* noise,
                             const float* logq, float* logq_out, float* kappa_out, float* xstats_out,
                             int B, int D, int M, float alpha, float alpha_bar, float beta,
                             uint64_t seed, int64_t sample_offset, int draw_index,
                             void* workspace, size_t workspace_bytes, void* stream);

/* ---- Self-attention core (north-star-only extension, NO reference code -- the reference
 * UNet, unet.py:37-65, has no attention; oracle = this repo's oracle.attention_core) ----
 * out[bh,q,:] = softmax_k(scale * <q[bh,q,:], k[bh,k,:]>) @ v[bh,k,:]   fused flash-style on tcgen05 / TMEM / TMA:
 * the S x S scores never leave the SM.  All tensors fp16, device, contiguous:
 *   q, k, out: [BH][S][64];   vt: V TRANSPOSED, [BH][64][S].   head_dim == 64, S % 128 == 0 (32^2 = 1024, 16^2 = 256). */
int sdd_attention_fwd(const void* q, const void* k, const void* vt, void* out, int BH, int S, int head_dim,
                      float scale, void* stream);
/* Self-attention BLOCK (same extension; oracle = oracle.attention_block):
 *   out = x + W_o attention(q, k, v) + b_o,   [q | k | v] = GroupNorm(4, 128)(x) W_qkv^T + b_qkv,   2 heads of 64
 * x, out: fp16 NHWC [B][S = H*W][128] (out may not alias x); gn_gamma / gn_beta fp32 [128]; w_qkv fp32 [384][128] (rows:
 * q, k, v; within each, head h = rows 64h..64h+63), b_qkv [384]; w_out fp32 [128][128], b_out [128].  S % 128 == 0.
 * GroupNorm statistics kernel -> projection GEMM (mma.sync, GroupNorm affine fused on load, head-split / V-transposed
 * epilogue) -> sdd_attention_fwd -> projection GEMM (+ bias + residual).  Synchronises the stream (scratch is freed). */
int sdd_attention_block_nhwc(const void* x, const float* gn_gamma, const float* gn_beta, const float* w_qkv,
                             const float* b_qkv, const float* w_out, const float* b_out, void* out, int B, int S, int C,
                             int heads, void* stream);
/* `iters` back-to-back launches between one CUDA-event pair; *ms_host = mean launch duration. */
int sdd_attention_profile(const void* q, const void* k, const void* vt, void* out, int BH, int S, int head_dim,
                          float scale, int iters, float* ms_host, void* stream);

/* ---- Training-side reuse: forward noising and the eps-MSE, forward only ----
 * out[b,:] = sqrt_ab[b] * x_start[b,:] + sqrt_1mab[b] * noise[b,:]                         (ddpm.py:13-17)
 * sqrt_ab / sqrt_1mab: device fp32[B], = torch.sqrt(alpha_bar[t_b]) and torch.sqrt(1 - alpha_bar[t_b]).
 * Same expression tree as the reference (no FMA contraction): bit-identical for the same inputs. */
int sdd_q_sample(const float* x_start, const float* noise, const float* sqrt_ab, const float* sqrt_1mab,
                 float* out, int B, int D, void* stream);
/* *out (device fp32 scalar) = mean((pred - target)^2) over n elements = F.mse_loss(pred, target), ddpm.py:24.
 * Deterministic (fixed-order double-precision partial sums in `workspace`, sdd_mse_workspace() bytes). */
size_t sdd_mse_workspace(void);
int sdd_mse(const float* pred, const float* target, size_t n, float* out, void* workspace,
            size_t workspace_bytes, void* stream);

/* Philox standard normals, same definition as the in-kernel noise (for x_T and for tests). */
int sdd_philox_normal(float* out, int B, int D, uint64_t seed, int64_t sample_offset, int draw_index,
                      void* stream);

/* ---- Sampler: the whole reverse loop (ddpm.py:31-45 for M = 1; superposed for M >= 2) ---- */
typedef struct {
  const float* noise_stack; /* device [T,B,D] (stack[0] = x_T, stack[k] = k-th in-loop draw) or NULL = Philox */
  uint64_t seed;            /* Philox key when noise_stack == NULL */
  int64_t sample_offset;    /* global index of local sample 0 (batch sharding) */
  float temperature;        /* 1.0 */
  const float* bias;        /* device [M] or NULL */
  float* x_out;             /* device [B,D] */
  float* kappa_traj;        /* device [T,B,M] or NULL */
  float* logq_traj;         /* device [T+1,B,M] or NULL */
  float* x_traj;            /* device [T+1,B,D] or NULL: row 0 = x_T, row k+1 = state after loop iteration k (the
                               reverse-diffusion strip of utils/visualization.py:6-28; per-step parity tests) */
  int use_graph;            /* 1: replay one captured step graph T-1 times (step 0 runs eagerly) */
  int mode;                 /* 0 = SuperDiff OR (kappa = softmax of the running log q); 1 = AND (kappa solved per
                               sample and step so that all models' log-density increments are equal; 8(f) N3) */
  const float* noise_host;  /* ABI 2: PINNED HOST [T,B,D] stack (same meaning as noise_stack; at most one of the two):
                               streamed to the device in step-range chunks on a copy stream while earlier steps compute,
                               through a two-chunk device ring -- O(chunk) device memory instead of the [T,B,D] stack,
                               and the host->device copy overlaps the loop instead of preceding it (ddpm.py:36 draws
                               one z per step; this is the explicit-noise equivalent of that O(B*D) footprint).  The
                               call returns with the copies still enqueued: the stack must stay valid and unmodified
                               until the work this call put on `stream` has completed */
  int noise_host_chunk;     /* slices per chunk of the streamed path; 0 = auto (~64 MB per chunk, at least one slice) */
} sdd_sample_args;

/* One step of the loop = the M UNet forwards (parallel branches of the captured step graph) + ONE update launch.  Inside the
 * step graph the update launch defers a sample's finalisation (log q increment, kappa, GroupNorm(1,1) statistics of x') to
 * the next step's consumers -- its own successor's prologue and the next forward's first layer -- so that the launch has
 * no serial tail; results are bit-identical to driving sdd_unet_forward_xstats + sdd_superpose_update step by step.
 * models[M] are borrowed and must outlive the sampler.  alphas/alpha_bars/betas are HOST fp32[T]
 * (ddpm.py:9-11).  All workspaces for (B,H,W) are allocated here. */
int sdd_sampler_create(sdd_sampler_t** out, sdd_unet_t* const* models, int M,
                       const float* alphas_host, const float* alpha_bars_host, const float* betas_host,
                       int T, int B, int H, int W, void* stream);
int sdd_sampler_run(sdd_sampler_t* s, const sdd_sample_args* args, void* stream);
int sdd_sampler_destroy(sdd_sampler_t* s);
/* Kernel launches one sdd_sampler_run issues (for bench.py's gpu_launches). */
int64_t sdd_sampler_launches_per_run(const sdd_sampler_t* s);
/* How many times this sampler captured + instantiated its step graph so far.  Seed, shard offset, noise stack,
 * temperature, bias and trajectory buffers live in device memory the graph points at, so this stays 1 across calls
 * (a change of `mode` or a re-allocated UNet workspace re-captures). */
int64_t sdd_sampler_graph_instantiations(const sdd_sampler_t* s);

/* ---- Operator-level entry points (used by the parity tests and the roofline bench) ---- */

/* The product conv kernel as an operator: GroupNorm(4,Cin)+SiLU of the RAW input (statistics in_meanrstd[B,4,2], affine
 * in_gamma/in_beta[Cin]) fused with the 3x3 conv (pad 1), bias add (bias[b*bias_batch_stride + c]) and the OUTPUT's
 * GroupNorm(4,Cout) statistics gn_meanrstd[B,4,2] (may be NULL) -- unet.py:21-28 in one launch.
 * act_raw / out: fp16 NHWC [B,H,W,Cin] / [B,H,W,Cout]; w: fp32 [Cout,Cin,3,3] (reference layout; rounded to fp16
 * internally).  Cin, Cout in {64,128}.  in_meanrstd == NULL: the input is used as is.  2-CTA tcgen05 kernel with
 * resident weights. */
int sdd_conv3x3_fused_nhwc(const void* act_raw, const float* in_meanrstd, const float* in_gamma, const float* in_beta,
                           const float* w, const float* bias, int64_t bias_batch_stride, void* out,
                           float* gn_meanrstd, int B, int H, int W, int Cin, int Cout, void* stream);

/* Kernel-only timing for the roofline numbers (bench.py): `iters` launches, each bracketed by CUDA events
 * on the launching stream; `flush` (may be NULL) is rewritten before every launch to evict L2.
 * *ms_host receives the mean kernel duration in milliseconds. */
/* impl: 1 = product kernel without the fused GroupNorm, 2 = product kernel with the fused GroupNorm+SiLU (identity
 * statistics); act / out fp16 NHWC. */
int sdd_conv3x3_profile(const void* act, const float* w, const float* bias, void* out, int B, int H, int W,
                        int Cin, int Cout, int impl, int iters, void* flush, size_t flush_bytes, float* ms_host,
                        void* stream);
/* Times the whole update step (one launch: HBM pass + per-sample finalize + step-counter bump), one event pair per
 * launch, `flush` rewritten before each. */
int sdd_superpose_update_profile(float* x, const float* eps, const float* noise, float* logq, int B, int D,
                                 int M, int iters, void* flush, size_t flush_bytes, float* ms_host,
                                 void* stream);

/* The update step's launch exactly as the sampler's step graph issues it (schedule table + device step counter,
 * deferred finalisation), timed as `iters` back-to-back launches between ONE event pair, each launch on the next of
 * ceil(rot_bytes / set bytes) (>= 2) private buffer sets (x, eps[M], optional noise), so that with rot_bytes >= 2 x L2 every launch
 * reads and writes data that is not in L2 while code, constants and TLBs stay warm (an L2 flush before every launch
 * also evicts the kernel's instructions, which dominates a 10-20 us kernel).  *ms_host = mean launch duration. */
int sdd_superpose_update_profile_rotating(int B, int D, int M, int use_noise, int iters, size_t rot_bytes,
                                          float* ms_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDD_B200_H */
