// C ABI (include/sdd_b200.h): handles, workspaces, launch orchestration, step-graph capture.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "conv_common.cuh"
#include "conv_tc4.cuh"
#include "unet_kernels.cuh"
#include "update.cuh"
#include "train_eval.cuh"
#include "attention.cuh"
#include "attention_block.cuh"

namespace sdd {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

static thread_local int64_t g_launches = 0;  // kernels launched by this thread since last reset
#define SDD_LAUNCH_CHECK()                 \
  do {                                     \
    ++::sdd::g_launches;                   \
    SDD_CUDA(cudaGetLastError());          \
  } while (0)

// ------------------------------------------------------------------------------- per-device state
// Function attributes (the opt-in to > 48 KB of dynamic shared memory) and the SM count are properties of a DEVICE, not
// of the process: they are set / queried once per device ordinal, so models on cuda:0 and cuda:1 of one process both work.
struct DeviceCtx {
  bool known = false;   // compute capability queried
  bool attrs = false;   // cudaFuncSetAttribute done on this device
  int major = 0, minor = 0, sms = 0;
};
static DeviceCtx g_dev[64];
static std::mutex g_dev_mu;

static int device_ctx(DeviceCtx** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error(std::string("no CUDA device: ") + cudaGetErrorString(e));
    cudaGetLastError();
    return SDD_ENODEV;
  }
  if (dev < 0 || dev >= 64) { set_error("device ordinal out of range"); return SDD_ENODEV; }
  std::lock_guard<std::mutex> lock(g_dev_mu);
  DeviceCtx& c = g_dev[dev];
  if (!c.known) {
    if (cudaDeviceGetAttribute(&c.major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&c.minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      set_error("cudaDeviceGetAttribute failed");
      cudaGetLastError();
      return SDD_ENODEV;
    }
    c.known = true;
  }
  if (c.major != 10) {
    set_error("device is sm_" + std::to_string(c.major) + std::to_string(c.minor) +
              "; this library contains sm_100a code only (no fallback)");
    return SDD_ENODEV;
  }
  if (out) *out = &c;
  return SDD_OK;
}
static int device_check() { return device_ctx(nullptr); }
static int num_sms() {
  DeviceCtx* c = nullptr;
  return device_ctx(&c) == SDD_OK ? c->sms : 1;
}

// ------------------------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}
// act fp16 [N][H][W][C]  ->  halo box (64 c, 10 w, 18 h, 1 n), 128-byte swizzle, zero fill out of bounds
static int make_act_map(CUtensorMap* m, const void* base, int N, int H, int W, int C) {
  EncodeTiledFn enc = get_encode();
  SDD_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)kHaloW, (cuuint32_t)(kTileH + 2), 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, SDD_ACT_TMAP_TYPE, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(act) failed: " + std::to_string((int)r));
    return SDD_ECUDA;
  }
  return SDD_OK;
}
// act fp16 [N][H][W][64] -> conv_out1's halo box (64 c, 34 w, 10 h, 1 n), 128-byte swizzle, zero fill out of bounds
static int make_o1_map(CUtensorMap* m, const void* base, int N, int H, int W) {
  EncodeTiledFn enc = get_encode();
  SDD_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
  cuuint32_t box[4] = {64, (cuuint32_t)kO1HW, (cuuint32_t)kO1HH, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, SDD_ACT_TMAP_TYPE, 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(conv_out1) failed: " + std::to_string((int)r));
    return SDD_ECUDA;
  }
  return SDD_OK;
}

// wt fp16 [9 = kx*3+ky][Cout][Cin] -> box (64 ci, Cout/2 rows, 1 tap): each CTA of a pair keeps its N half resident
static int make_wt_map(CUtensorMap* m, const void* base, int Cout, int Cin) {
  EncodeTiledFn enc = get_encode();
  SDD_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, 9};
  cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cout * Cin * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)(Cout / 2), 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, SDD_ACT_TMAP_TYPE, 3, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(wt) failed: " + std::to_string((int)r));
    return SDD_ECUDA;
  }
  return SDD_OK;
}

// fp16 [BH][rows][cols] (cols contiguous) -> box (64 cols = 128 B, box_rows, 1), 128-byte swizzle
static int make_attn_map(CUtensorMap* m, const void* base, int BH, int rows, int cols, int box_rows) {
  EncodeTiledFn enc = get_encode();
  SDD_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)BH};
  cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)rows * cols * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, SDD_ACT_TMAP_TYPE, 3, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(attention) failed: " + std::to_string((int)r));
    return SDD_ECUDA;
  }
  return SDD_OK;
}

// ------------------------------------------------------------------------------- conv launcher
static int ensure_func_attrs() {
  DeviceCtx* c = nullptr;
  SDD_TRY(device_ctx(&c));
  std::lock_guard<std::mutex> lock(g_dev_mu);
  if (c->attrs) return SDD_OK;
  const int reg_smem = conv_tc3_smem_bytes(128, 128, conv_tc3_stages(128, 128));
  // the four instantiations the launch policy uses: <Cout, Cin, raw ring>
  SDD_CUDA(cudaFuncSetAttribute(conv3x3_tc4_kernel<64, 64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, reg_smem));
  SDD_CUDA(cudaFuncSetAttribute(conv3x3_tc4_kernel<128, 128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, reg_smem));
  SDD_CUDA(cudaFuncSetAttribute(conv3x3_tc4_kernel<128, 64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kC2SmemLimit - 4096));
  SDD_CUDA(cudaFuncSetAttribute(conv3x3_tc4_kernel<64, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kC2SmemLimit - 4096));
  SDD_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
  SDD_CUDA(cudaFuncSetAttribute(conv_out1_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kO1SmemBytes));
  c->attrs = true;
  return SDD_OK;
}

struct GnInput3 {  // GroupNorm(4, Cin) + SiLU of the conv's input: fixed-point sums (product path) or mean/rstd floats
  const long long* sums;
  const float* meanrstd;
  const float* gamma;
  const float* beta;
  float* ab = nullptr;  // scratch [2][B][Cin] for gn_scale_shift_kernel
};

// Launch policy, from same-box A/B runs:
//  * generation 5 (kRaw: TMA-filled raw ring behind 3 operand stages) for the two Cin != Cout layers (64->128 314 -> 287,
//    128->64 369 -> 348 us per 32 x 256^2 chunk); 64->64 is slower with it (227 -> 244) and 128->128 has no room;
//  * contiguous tile ranges for 64->64 only (382 -> 373 us per 64-sample chunk; 64->128 508 -> 515): compile-time in the kernel;
//  * TMA L2 prefetch of a loader group's item after next on the register path (+3..6 %).
static int launch_conv(const CUtensorMap& tmA_halo, const CUtensorMap& tmB, const act_t* in, act_t* out, BiasRef bias,
                       GnInput3 gi, long long* out_sums, int B, int H, int W, int Cin, int Cout, cudaStream_t st) {
  SDD_CHECK(H % kTileH == 0 && W % kTileW == 0, "tcgen05 conv needs H % 16 == 0 and W % 8 == 0");
  SDD_CHECK((Cin == 64 || Cin == 128) && (Cout == 64 || Cout == 128), "tcgen05 conv supports 64/128 channels");
  SDD_CHECK((size_t)B * H * W * Cin * 2 < ((size_t)1 << 32), "input tensor of one launch must be < 4 GB (32-bit offsets)");
  SDD_TRY(ensure_func_attrs());
  ConvTc3Args a;
  a.in = in; a.out = out; a.bias = bias;
  a.in_ab = nullptr;
  if (gi.sums || gi.meanrstd) {
    SDD_CHECK(gi.ab && gi.gamma && gi.beta, "fused GroupNorm needs gamma, beta and the scale/shift scratch");
    gn_scale_shift_kernel<<<B, Cin, 0, st>>>(gi.sums, gi.meanrstd, gi.gamma, gi.beta,
                                             (double)H * (double)W * (double)(Cin / 4), kGnEps, gi.ab, B, Cin);
    SDD_LAUNCH_CHECK();
    a.in_ab = gi.ab;
  }
  a.out_sums = out_sums;
  a.B = B; a.H = H; a.W = W;
  a.tiles_w = W / kTileW;
  a.tiles_per_sample = (H / kTileH) * a.tiles_w;
  a.num_tiles = B * a.tiles_per_sample;
  a.num_pairs = (a.num_tiles + 1) / 2;
  a.stages = conv_tc3_stages(Cout, Cin);
  a.raw_slots = 0;
  a.prefetch = 1;
  if (Cin != Cout) {  // raw ring: 3 operand stages + up to 4 raw slots (both layers fit all 4)
    constexpr int kRawStages = 3, kRawMaxSlots = 4;
    int slots = kRawMaxSlots;
    while (slots > 3 && conv_tc3_smem_bytes(Cout, Cin, kRawStages + slots) > kC2SmemLimit - 4096) --slots;
    a.stages = kRawStages; a.raw_slots = slots;
  }
  const int smem = conv_tc3_smem_bytes(Cout, Cin, a.stages + a.raw_slots);
  SDD_CHECK(smem <= kC2SmemLimit - 4096, "conv shared-memory plan does not fit");
  const int grid = 2 * std::min(a.num_pairs, num_sms() / 2);
  if (Cin == 64 && Cout == 64) conv3x3_tc4_kernel<64, 64, false><<<grid, kC3Threads, smem, st>>>(tmA_halo, tmB, a);
  else if (Cin == 128 && Cout == 128) conv3x3_tc4_kernel<128, 128, false><<<grid, kC3Threads, smem, st>>>(tmA_halo, tmB, a);
  else if (Cin == 64) conv3x3_tc4_kernel<128, 64, true><<<grid, kC3Threads, smem, st>>>(tmA_halo, tmB, a);
  else conv3x3_tc4_kernel<64, 128, true><<<grid, kC3Threads, smem, st>>>(tmA_halo, tmB, a);
  SDD_LAUNCH_CHECK();
  return SDD_OK;
}

}  // namespace sdd

using namespace sdd;

// =============================================================================== UNet handle
namespace {

struct BlockParams {
  int cin, cout;
  const float *gn1_w, *gn1_b, *conv1_w, *conv1_b, *gn2_w, *gn2_b, *conv2_w, *conv2_b, *temb_w, *temb_b;
  act_t *conv1_wt, *conv2_wt;    // fp16 [kx][ky][Cout][Cin] (tensor-core convs only)
  CUtensorMap tm_w1h, tm_w2h;    // box (64 ci, Cout/2 rows, 1 tap)
};

constexpr int kBiasRow = 388;  // 64+128+128+64+1 = 385 per-block (conv2 bias + time_emb) values, padded to 16 B
constexpr int kBiasOff[5] = {0, 64, 192, 320, 384};
constexpr int kBlkCin[5] = {1, 64, 128, 128, 64};
constexpr int kBlkCout[5] = {64, 128, 128, 64, 1};
constexpr int kTimeDim = 256;
constexpr int kMaxBlocks = 10;
// ---- extension N2 (no reference code; oracle/unet_attn_oracle.py): multi-resolution UNet with attention at R/8 and R/16
// residual blocks: enc0 (1->64 @R), enc1 (64->128 @R/2), enc2..enc4 (128->128 @R/4, R/8, R/16), mid (128->128 @R/16),
// dec0 (128->128 @R/8), dec1 (128->128 @R/4), dec2 (128->64 @R/2), out (64->1 @R)
constexpr int kAnBlocks = 10;
constexpr int kAnCin[kAnBlocks] = {1, 64, 128, 128, 128, 128, 128, 128, 128, 64};
constexpr int kAnCout[kAnBlocks] = {64, 128, 128, 128, 128, 128, 128, 128, 64, 1};
constexpr int kAnLevel[kAnBlocks] = {0, 1, 2, 3, 4, 4, 3, 2, 1, 0};
constexpr int kAnBiasOff[kAnBlocks] = {0, 64, 192, 320, 448, 576, 704, 832, 960, 1024};
constexpr int kAnBiasRow = 1028;  // 1025 values, padded to 16 B
constexpr int kAnAttn = 4;        // attention blocks after enc3, enc4, mid, dec0
constexpr int kAnTensors = 30;    // activation tensors of one forward (see attn_forward_impl)
constexpr int kGnLayers = 9;  // GroupNorms fed by a conv output: downs.0 gn2 ... ups.1 gn2 (the first one reads x)

struct Workspace {
  int cap_b = 0, H = 0, W = 0;  // chunk capacity and image size the buffers were built for
  int64_t generation = 0;       // bumped on every (re)allocation; captured graphs check it
  act_t* act[2] = {nullptr, nullptr};
  float* e1 = nullptr;
  float* partials = nullptr;  // stats_x_kernel only
  int* counters = nullptr;
  long long* gnsums = nullptr;  // 9 x [cap_b][4][2] fixed-point GroupNorm sums, one slab per GroupNorm layer (common.cuh)
  float* xstats = nullptr;    // [cap_b][2] (used when the caller has no stats of x)
  float* gn_ab = nullptr;     // [2][cap_b][128] fused GroupNorm+SiLU scale / shift of the layer being launched
  CUtensorMap tm_halo[2][2];  // [buffer][Cin == 128], halo box (64, 10, 18)
  CUtensorMap tm_o1[2];       // [buffer] as 64-channel input of conv_out1: box (64, 34, 10)
  void release() {
    cudaFree(act[0]); cudaFree(act[1]); cudaFree(e1); cudaFree(partials); cudaFree(counters);
    cudaFree(gnsums); cudaFree(xstats); cudaFree(gn_ab);
    int64_t g = generation;
    *this = Workspace();
    generation = g;
  }
};

}  // namespace

namespace {
struct AttnParams { const float *gn_w, *gn_b, *w_qkv, *b_qkv, *w_out, *b_out; };
struct AnTensor {  // one activation tensor of the attention variant's forward
  act_t* p = nullptr;
  int level = 0, C = 0;
  long long* sums = nullptr;   // [cap_b][4][2] GroupNorm sums of this tensor
  CUtensorMap tm;              // halo map (conv inputs)
};

struct AttnWorkspace {
  int cap_b = 0, H = 0, W = 0;
  int64_t generation = 0;
  char* arena = nullptr;
  AnTensor t[kAnTensors];
  long long* sums_base = nullptr; size_t sums_bytes = 0;
  CUtensorMap tm_o1;           // tensor 29 as the input of conv_out1
  act_t *q = nullptr, *k = nullptr, *vt = nullptr, *ao = nullptr;  // attention scratch at the R/8 level's size
  float *e1 = nullptr, *gn_ab = nullptr, *xstats = nullptr, *partials = nullptr;
  int* counters = nullptr;
  void release() {
    cudaFree(arena);
    int64_t g = generation;
    *this = AttnWorkspace();
    generation = g;
  }
};
}  // namespace

struct sdd_unet {
  int arch = 0;             // 0 = the reference UNet (unet.py:37-65); 1 = multi-resolution attention variant (extension N2)
  int nblk = 5;             // residual blocks
  int bias_row = kBiasRow;  // floats per (conv2 bias + time embedding) row
  int bias_off[kMaxBlocks] = {0, 64, 192, 320, 384, 0, 0, 0, 0, 0};
  float* params = nullptr;  // one arena holding all fp32 tensors of the state_dict
  act_t* wt = nullptr;
  float* freq = nullptr;    // [128]
  int max_chunk = 0;        // sdd_unet_set_max_chunk: 0 = automatic
  const float *time_w1, *time_b1, *time_w2, *time_b2;
  BlockParams blk[kMaxBlocks];
  Workspace ws;
  // arch 1
  const float* class_emb = nullptr;  // [num_classes][256], added to the time embedding
  int num_classes = 0, label = 0;    // label: the class a sampler built on this handle conditions on
  AttnParams attn[kAnAttn];
  AttnWorkspace aws;
  // scratch for per-call time embeddings (forward with explicit t)
  int tscratch_n = 0;
  float *t_emb0 = nullptr, *t_h1 = nullptr, *t_emb = nullptr, *t_bias = nullptr;
};

namespace {

int attn_ensure_workspace(sdd_unet* u, int B, int H, int W);
int attn_forward_impl(sdd_unet* u, const float* x, XStatsSrc xsrc, BiasRef tb, float* eps_out, int B, int H, int W,
                      cudaStream_t st);

int chunk_for(const sdd_unet* u, int B, int H, int W) {
  if (u->max_chunk > 0) return std::min(u->max_chunk, B);
  // Measured (tools/conv_chunk.py): the conv kernels are far from HBM-bound (<= 1.5 TB/s of traffic), while every
  // launch pays a fixed prologue (resident-weight load, TMEM alloc) and a partially filled last wave, so large
  // chunks win: 128->128 goes from 0.53 of peak at 3 samples to 0.65 at >= 16.  Cap a ping-pong buffer at 1.5 GB.
  size_t per_sample = (size_t)H * W * 128 * 2;
  int c = (int)std::max<size_t>(1, ((size_t)1536 << 20) / per_sample);
  return std::min(c, B);
}

int ensure_workspace(sdd_unet* u, int B, int H, int W) {
  if (u->arch == 1) return attn_ensure_workspace(u, B, H, W);
  int need = chunk_for(u, B, H, W);
  Workspace& ws = u->ws;
  if (ws.cap_b >= need && ws.H == H && ws.W == W && (u->max_chunk == 0 || ws.cap_b == need)) return SDD_OK;
  ws.release();
  size_t act_bytes = (size_t)need * H * W * 128 * sizeof(act_t);
  SDD_CUDA(cudaMalloc(&ws.act[0], act_bytes));
  SDD_CUDA(cudaMalloc(&ws.act[1], act_bytes));
  SDD_CUDA(cudaMalloc(&ws.e1, (size_t)need * H * W * sizeof(float)));
  SDD_CUDA(cudaMalloc(&ws.partials, (size_t)need * kStatsBlocks * 2 * sizeof(float)));  // stats_x_kernel only
  SDD_CUDA(cudaMalloc(&ws.counters, (size_t)need * sizeof(int)));
  SDD_CUDA(cudaMemset(ws.counters, 0, (size_t)need * sizeof(int)));
  SDD_CUDA(cudaMalloc(&ws.gnsums, (size_t)kGnLayers * need * 8 * sizeof(long long)));
  SDD_CUDA(cudaMalloc(&ws.gn_ab, (size_t)2 * need * 128 * sizeof(float)));
  SDD_CUDA(cudaMalloc(&ws.xstats, (size_t)need * 2 * sizeof(float)));
  for (int bi = 0; bi < 2; ++bi) {
    SDD_TRY(make_act_map(&ws.tm_halo[bi][0], ws.act[bi], need, H, W, 64));
    SDD_TRY(make_act_map(&ws.tm_halo[bi][1], ws.act[bi], need, H, W, 128));
    SDD_TRY(make_o1_map(&ws.tm_o1[bi], ws.act[bi], need, H, W));
  }
  ws.cap_b = need; ws.H = H; ws.W = W;
  ++ws.generation;
  return SDD_OK;
}

__global__ void add_class_emb_kernel(float* emb, const float* table, const int64_t* y, int y_const, int num_classes) {
  const int i = blockIdx.x;
  int c = y ? (int)y[i] : y_const;
  c = c < 0 ? 0 : (c >= num_classes ? num_classes - 1 : c);
  emb[(size_t)i * kTimeDim + threadIdx.x] += table[(size_t)c * kTimeDim + threadIdx.x];
}

int64_t ws_generation(const sdd_unet* u) { return u->arch == 1 ? u->aws.generation : u->ws.generation; }

// rows[n][bias_row] = conv2.bias + time_emb(time_mlp(t_i) [+ class_emb(y_i)]) for every block (unet.py:33, :58)
int time_bias_rows(sdd_unet* u, const int64_t* t_dev, const int64_t* y_dev, int n, float* emb0, float* h1, float* emb,
                   float* rows, cudaStream_t st) {
  sinusoid_kernel<<<n, 128, 0, st>>>(t_dev, u->freq, emb0, n, kTimeDim / 2);
  SDD_LAUNCH_CHECK();
  auto lin = [&](const float* x, const float* Wm, const float* b, const float* extra, float* y, int K, int O,
                 int64_t ldy, int silu) {
    int warps = n * O;
    int blocks = (warps * 32 + 255) / 256;
    linear_kernel<<<blocks, 256, 0, st>>>(x, Wm, b, extra, y, n, K, O, ldy, silu);
  };
  lin(emb0, u->time_w1, u->time_b1, nullptr, h1, kTimeDim, 4 * kTimeDim, 4 * kTimeDim, 1);
  SDD_LAUNCH_CHECK();
  lin(h1, u->time_w2, u->time_b2, nullptr, emb, 4 * kTimeDim, kTimeDim, kTimeDim, 0);
  SDD_LAUNCH_CHECK();
  if (u->class_emb) {
    add_class_emb_kernel<<<n, kTimeDim, 0, st>>>(emb, u->class_emb, y_dev, u->label, u->num_classes);
    SDD_LAUNCH_CHECK();
  }
  for (int i = 0; i < u->nblk; ++i) {
    lin(emb, u->blk[i].temb_w, u->blk[i].temb_b, u->blk[i].conv2_b, rows + u->bias_off[i], kTimeDim, u->blk[i].cout,
        u->bias_row, 0);
    SDD_LAUNCH_CHECK();
  }
  return SDD_OK;
}

// One UNet forward over B samples, chunked so that a chunk's activations stay L2-resident.
// xsrc: where the first GroupNorm(1,1) gets (mean, rstd) of each x sample from (all null: computed here).
// tb: per-block bias rows; tb.base points at column 0 of the 385-wide row.
XStatsSrc xstats_src(const float* xstats) {
  XStatsSrc s;
  memset(&s, 0, sizeof(s));
  s.xstats = xstats;
  return s;
}
// the source for samples [b0, ...) of the batch; falls back to `computed` (statistics this forward computed itself)
XStatsSrc xstats_chunk(XStatsSrc s, int b0, const float* computed) {
  if (s.xstats) s.xstats += (size_t)b0 * 2;
  else if (s.partials) s.partials += (size_t)b0 * s.nblk * s.per_block;
  else s.xstats = computed;
  return s;
}

int unet_forward_impl(sdd_unet* u, const float* x, XStatsSrc xsrc, BiasRef tb, float* eps_out, int B, int H,
                      int W, cudaStream_t st) {
  if (u->arch == 1) return attn_forward_impl(u, x, xsrc, tb, eps_out, B, H, W, st);
  SDD_CHECK(H % kTileH == 0 && W % kTileW == 0, "H must be a multiple of 16 and W a multiple of 8");
  Workspace& ws = u->ws;
  SDD_CHECK(ws.cap_b >= 1 && ws.H == H && ws.W == W, "workspace not prepared");
  const int HW = H * W;
  const int chunk = ws.cap_b;
  const dim3 egrid((W + kCinTW - 1) / kCinTW, (H + kCinTH - 1) / kCinTH, 1);
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int nb = std::min(chunk, B - b0);
    const float* xc = x + (size_t)b0 * HW;
    auto sums = [&](int i) { return ws.gnsums + (size_t)i * ws.cap_b * 8; };
    auto bias_const = [&](const float* p) { return BiasRef{p, nullptr, 0, 0}; };
    auto bias_time = [&](int blk) {
      return BiasRef{tb.base + kBiasOff[blk] + (int64_t)b0 * tb.batch_stride, tb.row_ptr, tb.row_stride,
                     tb.batch_stride};
    };
    // every GroupNorm accumulator of this forward starts at zero (producers only ever RED.ADD into them)
    SDD_CUDA(cudaMemsetAsync(ws.gnsums, 0, (size_t)kGnLayers * ws.cap_b * 8 * sizeof(long long), st));
    if (!xsrc.xstats && !xsrc.partials) {
      stats_x_kernel<<<dim3(kStatsBlocks, nb), 256, 0, st>>>(xc, HW, ws.partials, ws.counters, ws.xstats);
      SDD_LAUNCH_CHECK();
    }
    const XStatsSrc xs = xstats_chunk(xsrc, b0, ws.xstats);
    dim3 eg = egrid; eg.z = nb;
    // downs.0: GN(1,1)+SiLU fused into the 1->64 conv; raw result in act[0], its GroupNorm(4,64) sums -> sums(0)
    const BlockParams& d0 = u->blk[0];
    {
      const int in_tiles = (int)(eg.x * eg.y * eg.z);
      conv_in_mma_kernel<<<std::min(in_tiles, 2 * num_sms()), 256, 0, st>>>(  // two persistent CTAs per SM
          xc, xs, d0.gn1_w, d0.gn1_b, d0.conv1_w, bias_const(d0.conv1_b), ws.act[0], sums(0), H, W, (int)eg.x, (int)eg.y,
          in_tiles);
    }
    SDD_LAUNCH_CHECK();
    // every tensor-core conv normalises + activates its own input (GroupNorm+SiLU fused on the operand path)
    SDD_TRY(launch_conv(ws.tm_halo[0][0], d0.tm_w2h, ws.act[0], ws.act[1], bias_time(0),
                            GnInput3{sums(0), nullptr, d0.gn2_w, d0.gn2_b, ws.gn_ab}, sums(1), nb, H, W, 64, 64, st));
    int cur = 1, gi = 1;
    for (int bi = 1; bi <= 3; ++bi) {
      const BlockParams& p = u->blk[bi];
      SDD_TRY(launch_conv(ws.tm_halo[cur][p.cin == 128], p.tm_w1h, ws.act[cur], ws.act[cur ^ 1], bias_const(p.conv1_b),
                              GnInput3{sums(gi), nullptr, p.gn1_w, p.gn1_b, ws.gn_ab}, sums(gi + 1), nb, H, W, p.cin, p.cout, st));
      cur ^= 1; ++gi;
      SDD_TRY(launch_conv(ws.tm_halo[cur][p.cout == 128], p.tm_w2h, ws.act[cur], ws.act[cur ^ 1], bias_time(bi),
                              GnInput3{sums(gi), nullptr, p.gn2_w, p.gn2_b, ws.gn_ab}, sums(gi + 1), nb, H, W, p.cout, p.cout, st));
      cur ^= 1; ++gi;
    }
    // ups.1: GroupNorm(4,64)+SiLU fused into the 64->1 conv (mma.sync), then GN(1,1)+SiLU fused into the 1->1 conv
    const BlockParams& u1 = u->blk[4];
    {
      const int o1_tiles = (int)(eg.x * eg.y * eg.z);
      SDD_TRY(ensure_func_attrs());
      conv_out1_tma_kernel<<<std::min(o1_tiles, 2 * num_sms()), 256, kO1SmemBytes, st>>>(  // two persistent CTAs per SM
          ws.tm_o1[cur], sums(gi), u1.gn1_w, u1.gn1_b, u1.conv1_w, u1.conv1_b, ws.e1, sums(gi + 1), H, W, (int)eg.x,
          (int)eg.y, o1_tiles);
    }
    SDD_LAUNCH_CHECK();
    conv_out2_kernel<<<eg, 256, 0, st>>>(ws.e1, sums(gi + 1), u1.gn2_w, u1.gn2_b, u1.conv2_w, bias_time(4),
                                         eps_out + (size_t)b0 * HW, H, W);
    SDD_LAUNCH_CHECK();
  }
  return SDD_OK;
}


// ------------------------------------------------------------------------------- attention block (graph-capturable)
// out = x + W_o attention(q, k, v) + b_o with [q | k | v] = GroupNorm(4,128)(x) W_qkv^T + b_qkv: three launches, scratch
// from the caller, no allocation and no synchronisation.  GroupNorm statistics of x: fixed-point sums from its producer
// (in_sums) or (mean, rstd) floats; out_sums (optional) receives the output's GroupNorm sums for the next layer.
int launch_attention_block(const act_t* x, const long long* in_sums, const float* in_meanrstd, const AttnParams& p,
                           act_t* q, act_t* k, act_t* vt, act_t* ao, act_t* out, long long* out_sums, int B, int S,
                           cudaStream_t st) {
  SDD_CHECK(B > 0 && S >= kAttnBN && S % kAttnBN == 0, "attention needs S = H*W a positive multiple of 128");
  SDD_CHECK(B * kAbHeads <= 65535, "batch * heads must be <= 65535");
  SDD_TRY(ensure_func_attrs());
  AttnBlockGemmArgs g;
  memset(&g, 0, sizeof(g));
  g.a = x; g.w = p.w_qkv; g.bias = p.b_qkv; g.in_sums = in_sums; g.meanrstd = in_meanrstd; g.gamma = p.gn_w; g.beta = p.gn_b;
  g.q = q; g.k = k; g.vt = vt; g.B = B; g.S = S;
  attn_block_gemm_kernel<0><<<dim3((unsigned)((size_t)B * S / kAbRows), 3 * kAbC / 64), 128, 0, st>>>(g);
  SDD_LAUNCH_CHECK();
  const int BH = B * kAbHeads;
  CUtensorMap tmQ, tmK, tmVt;
  SDD_TRY(make_attn_map(&tmQ, q, BH, S, kAttnD, kAttnBM));
  SDD_TRY(make_attn_map(&tmK, k, BH, S, kAttnD, kAttnBN));
  SDD_TRY(make_attn_map(&tmVt, vt, BH, kAttnD, S, kAttnD));
  AttnArgs a;
  a.out = ao; a.S = S; a.BH = BH;
  a.scale_log2e = 0.125f * 1.4426950408889634f;
  attention_fwd_kernel<<<dim3((unsigned)(S / kAttnBM), (unsigned)BH), kAttnThreads, kAttnSmem, st>>>(tmQ, tmK, tmVt, a);
  SDD_LAUNCH_CHECK();
  memset(&g, 0, sizeof(g));
  g.a = ao; g.w = p.w_out; g.bias = p.b_out; g.resid = x; g.out = out; g.out_sums = out_sums; g.B = B; g.S = S;
  attn_block_gemm_kernel<1><<<dim3((unsigned)((size_t)B * S / kAbRows), kAbC / 64), 128, 0, st>>>(g);
  SDD_LAUNCH_CHECK();
  return SDD_OK;
}

// ------------------------------------------------------------------------------- extension N2: attention-variant forward
constexpr int kAnTLevel[kAnTensors] = {0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 3, 3, 3, 3, 2, 2, 2, 1, 1, 1, 0};
constexpr int kAnTC[kAnTensors] = {64, 64, 64, 128, 128, 128, 128, 128, 128, 128, 128, 128, 128, 128, 128,
                                   128, 128, 128, 128, 128, 128, 128, 128, 128, 128, 128, 128, 64, 64, 64};

int attn_ensure_workspace(sdd_unet* u, int B, int H, int W) {
  SDD_CHECK(H % 256 == 0 && W % 128 == 0 && ((H >> 4) * (W >> 4)) % kAttnBN == 0,
            "the attention variant needs H % 256 == 0 and W % 128 == 0 (five levels, attention at R/8 and R/16)");
  // ~52 MB of activations per sample at 256^2: cap a chunk at ~4 GB
  const size_t per_sample = (size_t)H * W * 800;
  int need = u->max_chunk > 0 ? std::min(u->max_chunk, B) : (int)std::min<size_t>((size_t)B, std::max<size_t>(1, ((size_t)4 << 30) / per_sample));
  AttnWorkspace& ws = u->aws;
  if (ws.cap_b >= need && ws.H == H && ws.W == W && (u->max_chunk == 0 || ws.cap_b == need)) return SDD_OK;
  ws.release();
  auto al = [](size_t v) { return (v + 1023) & ~(size_t)1023; };
  size_t off = 0, toff[kAnTensors];
  for (int i = 0; i < kAnTensors; ++i) {
    toff[i] = off;
    off += al((size_t)need * (H >> kAnTLevel[i]) * (W >> kAnTLevel[i]) * kAnTC[i] * sizeof(act_t));
  }
  const size_t attn_elems = (size_t)need * (H >> 3) * (W >> 3) * kAbC;
  const size_t o_attn = off; off += al(4 * attn_elems * sizeof(act_t));
  const size_t o_e1 = off; off += al((size_t)need * H * W * sizeof(float));
  const size_t o_sums = off; const size_t sums_bytes = (size_t)(kAnTensors + 1) * need * 8 * sizeof(long long); off += al(sums_bytes);
  const size_t o_ab = off; off += al((size_t)2 * need * 128 * sizeof(float));
  const size_t o_xs = off; off += al((size_t)need * 2 * sizeof(float));
  const size_t o_part = off; off += al((size_t)need * kStatsBlocks * 2 * sizeof(float));
  const size_t o_cnt = off; off += al((size_t)need * sizeof(int));
  if (cudaMalloc(&ws.arena, off) != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc(attention-variant workspace) failed"); return SDD_ENOMEM; }
  ws.sums_base = reinterpret_cast<long long*>(ws.arena + o_sums); ws.sums_bytes = sums_bytes;
  for (int i = 0; i < kAnTensors; ++i) {
    AnTensor& t = ws.t[i];
    t.p = reinterpret_cast<act_t*>(ws.arena + toff[i]); t.level = kAnTLevel[i]; t.C = kAnTC[i];
    t.sums = ws.sums_base + (size_t)i * need * 8;
    SDD_TRY(make_act_map(&t.tm, t.p, need, H >> t.level, W >> t.level, t.C));
  }
  SDD_TRY(make_o1_map(&ws.tm_o1, ws.t[29].p, need, H, W));
  act_t* ab = reinterpret_cast<act_t*>(ws.arena + o_attn);
  ws.q = ab; ws.k = ab + attn_elems; ws.vt = ab + 2 * attn_elems; ws.ao = ab + 3 * attn_elems;
  ws.e1 = reinterpret_cast<float*>(ws.arena + o_e1);
  ws.gn_ab = reinterpret_cast<float*>(ws.arena + o_ab);
  ws.xstats = reinterpret_cast<float*>(ws.arena + o_xs);
  ws.partials = reinterpret_cast<float*>(ws.arena + o_part);
  ws.counters = reinterpret_cast<int*>(ws.arena + o_cnt);
  SDD_CUDA(cudaMemset(ws.counters, 0, (size_t)need * sizeof(int)));
  ws.cap_b = need; ws.H = H; ws.W = W;
  ++ws.generation;
  return SDD_OK;
}

int launch_resample(bool up, const AnTensor& in, const AnTensor* skip, const AnTensor& out, bool want_sums, int nb, int H,
                    int W, cudaStream_t st) {
  const int Ho = H >> out.level, Wo = W >> out.level, C = out.C;
  const int nvec = Ho * Wo * (C / 8);
  const int blocks = std::max(1, std::min((nvec + 255) / 256, 8 * num_sms()));  // a function of the shape only (never of B)
  if (up) resample_kernel<true><<<dim3(blocks, nb), 256, 0, st>>>(in.p, skip->p, out.p, want_sums ? out.sums : nullptr, Ho, Wo, C);
  else resample_kernel<false><<<dim3(blocks, nb), 256, 0, st>>>(in.p, nullptr, out.p, want_sums ? out.sums : nullptr, Ho, Wo, C);
  SDD_LAUNCH_CHECK();
  return SDD_OK;
}

// One forward of the attention variant over B samples (chunked).  Same contract as unet_forward_impl.
int attn_forward_impl(sdd_unet* u, const float* x, XStatsSrc xsrc, BiasRef tb, float* eps_out, int B, int H, int W,
                      cudaStream_t st) {
  AttnWorkspace& ws = u->aws;
  SDD_CHECK(ws.cap_b >= 1 && ws.H == H && ws.W == W, "workspace not prepared");
  const int HW = H * W;
  const dim3 egrid((W + kCinTW - 1) / kCinTW, (H + kCinTH - 1) / kCinTH, 1);
  for (int b0 = 0; b0 < B; b0 += ws.cap_b) {
    const int nb = std::min(ws.cap_b, B - b0);
    const float* xc = x + (size_t)b0 * HW;
    auto bias_const = [&](const float* p) { return BiasRef{p, nullptr, 0, 0}; };
    auto bias_time = [&](int blk) {
      return BiasRef{tb.base + u->bias_off[blk] + (int64_t)b0 * tb.batch_stride, tb.row_ptr, tb.row_stride, tb.batch_stride};
    };
    AnTensor* T = ws.t;
    SDD_CUDA(cudaMemsetAsync(ws.sums_base, 0, ws.sums_bytes, st));
    if (!xsrc.xstats && !xsrc.partials) {
      stats_x_kernel<<<dim3(kStatsBlocks, nb), 256, 0, st>>>(xc, HW, ws.partials, ws.counters, ws.xstats);
      SDD_LAUNCH_CHECK();
    }
    const XStatsSrc xs = xstats_chunk(xsrc, b0, ws.xstats);
    // conv of a residual block: input tensor `i` (GroupNorm+SiLU fused on the operand path) -> output tensor `o`
    auto conv = [&](int i, int o, const CUtensorMap& tmw, BiasRef bias, const float* gw, const float* gb, bool want_sums) {
      const AnTensor& ti = T[i];
      return launch_conv(ti.tm, tmw, ti.p, T[o].p, bias, GnInput3{ti.sums, nullptr, gw, gb, ws.gn_ab},
                         want_sums ? T[o].sums : nullptr, nb, H >> ti.level, W >> ti.level, ti.C, T[o].C, st);
    };
    auto rb = [&](int blk, int i, int mid, int o, bool want_sums) {  // blocks 1..8: two tensor-core convs
      const BlockParams& p = u->blk[blk];
      SDD_TRY(conv(i, mid, p.tm_w1h, bias_const(p.conv1_b), p.gn1_w, p.gn1_b, true));
      return conv(mid, o, p.tm_w2h, bias_time(blk), p.gn2_w, p.gn2_b, want_sums);
    };
    auto attn = [&](int ai, int i, int o, bool want_sums) {
      const int S = (H >> T[i].level) * (W >> T[i].level);
      return launch_attention_block(T[i].p, T[i].sums, nullptr, u->attn[ai], ws.q, ws.k, ws.vt, ws.ao, T[o].p,
                                    want_sums ? T[o].sums : nullptr, nb, S, st);
    };
    // ---- enc0 @R: 1 -> 64 (TF32 mma.sync), 64 -> 64
    dim3 eg = egrid; eg.z = nb;
    {
      const BlockParams& p = u->blk[0];
      const int in_tiles = (int)(eg.x * eg.y * eg.z);
      conv_in_mma_kernel<<<std::min(in_tiles, 2 * num_sms()), 256, 0, st>>>(
          xc, xs, p.gn1_w, p.gn1_b, p.conv1_w, bias_const(p.conv1_b), T[0].p, T[0].sums, H, W, (int)eg.x, (int)eg.y, in_tiles);
      SDD_LAUNCH_CHECK();
      SDD_TRY(conv(0, 1, p.tm_w2h, bias_time(0), p.gn2_w, p.gn2_b, false));
    }
    // ---- encoder: average-pool, residual block (+ attention at R/8 and R/16)
    SDD_TRY(launch_resample(false, T[1], nullptr, T[2], true, nb, H, W, st));
    SDD_TRY(rb(1, 2, 3, 4, false));
    SDD_TRY(launch_resample(false, T[4], nullptr, T[5], true, nb, H, W, st));
    SDD_TRY(rb(2, 5, 6, 7, false));
    SDD_TRY(launch_resample(false, T[7], nullptr, T[8], true, nb, H, W, st));
    SDD_TRY(rb(3, 8, 9, 10, true));
    SDD_TRY(attn(0, 10, 11, false));
    SDD_TRY(launch_resample(false, T[11], nullptr, T[12], true, nb, H, W, st));
    SDD_TRY(rb(4, 12, 13, 14, true));
    SDD_TRY(attn(1, 14, 15, true));
    // ---- middle @R/16
    SDD_TRY(rb(5, 15, 16, 17, true));
    SDD_TRY(attn(2, 17, 18, false));
    // ---- decoder: nearest-neighbour upsampling + additive skip, residual block (+ attention at R/8)
    SDD_TRY(launch_resample(true, T[18], &T[11], T[19], true, nb, H, W, st));
    SDD_TRY(rb(6, 19, 20, 21, true));
    SDD_TRY(attn(3, 21, 22, false));
    SDD_TRY(launch_resample(true, T[22], &T[7], T[23], true, nb, H, W, st));
    SDD_TRY(rb(7, 23, 24, 25, false));
    SDD_TRY(launch_resample(true, T[25], &T[4], T[26], true, nb, H, W, st));
    SDD_TRY(rb(8, 26, 27, 28, false));
    SDD_TRY(launch_resample(true, T[28], &T[1], T[29], true, nb, H, W, st));
    // ---- out @R: 64 -> 1 (fp16 mma.sync), 1 -> 1
    {
      const BlockParams& p = u->blk[9];
      long long* e1_sums = ws.sums_base + (size_t)kAnTensors * ws.cap_b * 8;
      const int o1_tiles = (int)(eg.x * eg.y * eg.z);
      conv_out1_tma_kernel<<<std::min(o1_tiles, 2 * num_sms()), 256, kO1SmemBytes, st>>>(
          ws.tm_o1, T[29].sums, p.gn1_w, p.gn1_b, p.conv1_w, p.conv1_b, ws.e1, e1_sums, H, W, (int)eg.x, (int)eg.y, o1_tiles);
      SDD_LAUNCH_CHECK();
      conv_out2_kernel<<<eg, 256, 0, st>>>(ws.e1, e1_sums, p.gn2_w, p.gn2_b, p.conv2_w, bias_time(9),
                                           eps_out + (size_t)b0 * HW, H, W);
      SDD_LAUNCH_CHECK();
    }
  }
  return SDD_OK;
}


// fp16 kernel-layout copies + tensor maps of every tensor-core conv's weights, the sinusoid frequency table; syncs.
int build_conv_weights(sdd_unet* u, cudaStream_t st) {
  size_t wt_elems = 0;
  for (int i = 0; i < u->nblk; ++i) {
    const BlockParams& b = u->blk[i];
    if (b.cin >= 64 && b.cout >= 64) wt_elems += (size_t)9 * b.cin * b.cout;
    if (b.cout >= 64) wt_elems += (size_t)9 * b.cout * b.cout;
  }
  if (cudaMalloc(&u->wt, wt_elems * sizeof(act_t)) != cudaSuccess) { set_error("cudaMalloc(wt) failed"); return SDD_ENOMEM; }
  act_t* wp = u->wt;
  for (int i = 0; i < u->nblk; ++i) {
    BlockParams& b = u->blk[i];
    auto conv = [&](const float* w, int cout, int cin, act_t** dst, CUtensorMap* tmh) -> int {
      int total_w = 9 * cout * cin;
      conv_weight_to_act_kernel<<<(total_w + 255) / 256, 256, 0, st>>>(w, wp, cout, cin);
      SDD_LAUNCH_CHECK();
      *dst = wp;
      SDD_TRY(make_wt_map(tmh, wp, cout, cin));
      wp += total_w;
      return SDD_OK;
    };
    if (b.cin >= 64 && b.cout >= 64) SDD_TRY(conv(b.conv1_w, b.cout, b.cin, &b.conv1_wt, &b.tm_w1h));
    if (b.cout >= 64) SDD_TRY(conv(b.conv2_w, b.cout, b.cout, &b.conv2_wt, &b.tm_w2h));
  }
  // sinusoid frequencies exactly as unet.py:14 evaluates them in fp32
  float hf[kTimeDim / 2];
  const float step = -(std::log(10000.0f) / (float)(kTimeDim / 2 - 1));
  for (int k = 0; k < kTimeDim / 2; ++k) hf[k] = std::exp((float)k * step);
  if (cudaMalloc(&u->freq, sizeof(hf)) != cudaSuccess) { set_error("cudaMalloc(freq) failed"); return SDD_ENOMEM; }
  SDD_CUDA(cudaMemcpyAsync(u->freq, hf, sizeof(hf), cudaMemcpyHostToDevice, st));
  SDD_CUDA(cudaStreamSynchronize(st));
  return SDD_OK;
}

}  // namespace

// =============================================================================== C ABI
extern "C" {

int sdd_abi_version(void) { return SDD_ABI_VERSION; }
const char* sdd_last_error(void) { return g_err.c_str(); }
int sdd_device_check(void) { return device_check(); }

int sdd_unet_create(sdd_unet_t** out, const float* const* tensors, int num_tensors, void* stream) {
  SDD_CHECK(out && tensors, "null argument");
  SDD_CHECK(num_tensors == SDD_UNET_NUM_TENSORS, "expected the 54 tensors of the reference UNet state_dict");
  SDD_TRY(device_check());
  cudaStream_t st = (cudaStream_t)stream;
  // element counts in state_dict order
  std::vector<size_t> n;
  n.insert(n.end(), {(size_t)1024 * 256, 1024, (size_t)256 * 1024, 256});
  for (int i = 0; i < 5; ++i) {
    size_t ci = kBlkCin[i], co = kBlkCout[i];
    n.insert(n.end(), {ci, ci, co * ci * 9, co, co, co, co * co * 9, co, co * 256, co});
  }
  size_t total = 0;
  std::vector<size_t> off(n.size());
  for (size_t i = 0; i < n.size(); ++i) { off[i] = total; total += (n[i] + 3) & ~(size_t)3; }
  sdd_unet* u = new sdd_unet();
  auto fail = [&](int code) { sdd_unet_destroy(u); return code; };
  if (cudaMalloc(&u->params, total * sizeof(float)) != cudaSuccess) { set_error("cudaMalloc(params) failed"); return fail(SDD_ENOMEM); }
  for (size_t i = 0; i < n.size(); ++i) {
    if (cudaMemcpyAsync(u->params + off[i], tensors[i], n[i] * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
      set_error("copying state-dict tensor " + std::to_string(i) + " failed (device pointers expected)");
      return fail(SDD_ECUDA);
    }
  }
  const float* P = u->params;
  u->time_w1 = P + off[0]; u->time_b1 = P + off[1]; u->time_w2 = P + off[2]; u->time_b2 = P + off[3];
  for (int i = 0; i < 5; ++i) {
    BlockParams& b = u->blk[i];
    const size_t* o = &off[4 + 10 * i];
    b.cin = kBlkCin[i]; b.cout = kBlkCout[i];
    b.gn1_w = P + o[0]; b.gn1_b = P + o[1]; b.conv1_w = P + o[2]; b.conv1_b = P + o[3];
    b.gn2_w = P + o[4]; b.gn2_b = P + o[5]; b.conv2_w = P + o[6]; b.conv2_b = P + o[7];
    b.temb_w = P + o[8]; b.temb_b = P + o[9];
    b.conv1_wt = b.conv2_wt = nullptr;
  }
  int rc = build_conv_weights(u, st);
  if (rc != SDD_OK) return fail(rc);
  *out = u;
  return SDD_OK;
}

int sdd_unet_attn_create(sdd_unet_t** out, const float* const* tensors, int num_tensors, int num_classes, void* stream) {
  SDD_CHECK(out && tensors, "null argument");
  SDD_CHECK(num_classes >= 1, "num_classes must be >= 1");
  SDD_CHECK(num_tensors == SDD_UNET_ATTN_NUM_TENSORS, "expected the 129 tensors of the UNetAttn state_dict");
  SDD_TRY(device_check());
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<size_t> n;
  n.insert(n.end(), {(size_t)1024 * 256, 1024, (size_t)256 * 1024, 256, (size_t)num_classes * 256});
  for (int i = 0; i < kAnBlocks; ++i) {
    size_t ci = kAnCin[i], co = kAnCout[i];
    n.insert(n.end(), {ci, ci, co * ci * 9, co, co, co, co * co * 9, co, co * 256, co});
  }
  for (int i = 0; i < kAnAttn; ++i) n.insert(n.end(), {128, 128, (size_t)384 * 128, 384, (size_t)128 * 128, 128});
  size_t total = 0;
  std::vector<size_t> off(n.size());
  for (size_t i = 0; i < n.size(); ++i) { off[i] = total; total += (n[i] + 3) & ~(size_t)3; }
  sdd_unet* u = new sdd_unet();
  u->arch = 1; u->nblk = kAnBlocks; u->bias_row = kAnBiasRow; u->num_classes = num_classes;
  for (int i = 0; i < kAnBlocks; ++i) u->bias_off[i] = kAnBiasOff[i];
  auto fail = [&](int code) { sdd_unet_destroy(u); return code; };
  if (cudaMalloc(&u->params, total * sizeof(float)) != cudaSuccess) { set_error("cudaMalloc(params) failed"); return fail(SDD_ENOMEM); }
  for (size_t i = 0; i < n.size(); ++i) {
    if (cudaMemcpyAsync(u->params + off[i], tensors[i], n[i] * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
      set_error("copying state-dict tensor " + std::to_string(i) + " failed (device pointers expected)");
      return fail(SDD_ECUDA);
    }
  }
  const float* P = u->params;
  u->time_w1 = P + off[0]; u->time_b1 = P + off[1]; u->time_w2 = P + off[2]; u->time_b2 = P + off[3];
  u->class_emb = P + off[4];
  for (int i = 0; i < kAnBlocks; ++i) {
    BlockParams& b = u->blk[i];
    const size_t* o = &off[5 + 10 * i];
    b.cin = kAnCin[i]; b.cout = kAnCout[i];
    b.gn1_w = P + o[0]; b.gn1_b = P + o[1]; b.conv1_w = P + o[2]; b.conv1_b = P + o[3];
    b.gn2_w = P + o[4]; b.gn2_b = P + o[5]; b.conv2_w = P + o[6]; b.conv2_b = P + o[7];
    b.temb_w = P + o[8]; b.temb_b = P + o[9];
    b.conv1_wt = b.conv2_wt = nullptr;
  }
  for (int i = 0; i < kAnAttn; ++i) {
    const size_t* o = &off[5 + 10 * kAnBlocks + 6 * i];
    u->attn[i] = AttnParams{P + o[0], P + o[1], P + o[2], P + o[3], P + o[4], P + o[5]};
  }
  int rc = build_conv_weights(u, st);
  if (rc != SDD_OK) return fail(rc);
  *out = u;
  return SDD_OK;
}

int sdd_unet_set_label(sdd_unet_t* u, int label) {
  SDD_CHECK(u && u->arch == 1, "class labels need the class-conditional variant (sdd_unet_attn_create)");
  SDD_CHECK(label >= 0 && label < u->num_classes, "label out of range");
  u->label = label;
  return SDD_OK;
}

int sdd_unet_destroy(sdd_unet_t* u) {
  if (!u) return SDD_OK;
  u->ws.release();
  u->aws.release();
  cudaFree(u->params); cudaFree(u->wt); cudaFree(u->freq);
  cudaFree(u->t_emb0); cudaFree(u->t_h1); cudaFree(u->t_emb); cudaFree(u->t_bias);
  delete u;
  return SDD_OK;
}

int sdd_unet_set_max_chunk(sdd_unet_t* u, int max_samples) {
  SDD_CHECK(u && max_samples >= 0, "bad argument");
  u->max_chunk = max_samples;
  return SDD_OK;
}

int sdd_unet_forward(sdd_unet_t* u, const float* x, const int64_t* t, float* eps_out, int B, int H, int W,
                     void* stream) {
  return sdd_unet_forward_xstats(u, x, nullptr, t, eps_out, B, H, W, stream);
}

int sdd_unet_forward_xstats(sdd_unet_t* u, const float* x, const float* xstats, const int64_t* t, float* eps_out, int B,
                            int H, int W, void* stream) {
  return sdd_unet_forward_labeled(u, x, xstats, t, nullptr, eps_out, B, H, W, stream);
}

int sdd_unet_forward_labeled(sdd_unet_t* u, const float* x, const float* xstats, const int64_t* t, const int64_t* y,
                             float* eps_out, int B, int H, int W, void* stream) {
  SDD_CHECK(u && x && t && eps_out, "null argument");
  SDD_CHECK(!y || u->arch == 1, "class labels need the class-conditional variant (sdd_unet_attn_create)");
  SDD_CHECK(B >= 1 && H >= 16 && W >= 8, "bad shape");
  SDD_CHECK(H % kTileH == 0 && W % kTileW == 0, "H must be a multiple of 16 and W a multiple of 8");
  cudaStream_t st = (cudaStream_t)stream;
  SDD_TRY(ensure_workspace(u, B, H, W));
  if (u->tscratch_n < B) {
    cudaFree(u->t_emb0); cudaFree(u->t_h1); cudaFree(u->t_emb); cudaFree(u->t_bias);
    u->t_emb0 = u->t_h1 = u->t_emb = u->t_bias = nullptr; u->tscratch_n = 0;
    SDD_CUDA(cudaMalloc(&u->t_emb0, (size_t)B * kTimeDim * sizeof(float)));
    SDD_CUDA(cudaMalloc(&u->t_h1, (size_t)B * 4 * kTimeDim * sizeof(float)));
    SDD_CUDA(cudaMalloc(&u->t_emb, (size_t)B * kTimeDim * sizeof(float)));
    SDD_CUDA(cudaMalloc(&u->t_bias, (size_t)B * u->bias_row * sizeof(float)));
    u->tscratch_n = B;
  }
  SDD_TRY(time_bias_rows(u, t, y, B, u->t_emb0, u->t_h1, u->t_emb, u->t_bias, st));
  BiasRef tb{u->t_bias, nullptr, 0, u->bias_row};
  return unet_forward_impl(u, x, xstats_src(xstats), tb, eps_out, B, H, W, st);
}

// ------------------------------------------------------------------------------- fused update
size_t sdd_superpose_update_workspace(int B, int D, int M) { return update_workspace_bytes(B, D, M); }

}  // extern "C"

namespace sdd {

int launch_superpose_update(UpdateArgs a, void* workspace, cudaStream_t st) {
  SDD_CHECK(a.M >= 1 && a.M <= kMaxModels, "1 <= M <= 4");
  SDD_CHECK(a.D % 4 == 0 && a.D > 0 && a.B > 0, "D must be a positive multiple of 4");
  SDD_CHECK(a.mode == 0 || a.mode == 1, "mode must be 0 (OR) or 1 (AND)");
  char* w = reinterpret_cast<char*>(workspace);
  a.partials = reinterpret_cast<float*>(w);
  a.parity_stride = update_ws_part_bytes(a.B, a.D) / sizeof(float);
  w += 2 * update_ws_part_bytes(a.B, a.D);
  a.and_partials = reinterpret_cast<float*>(w);
  w += update_ws_and_bytes(a.B, a.D);
  a.kappa_in = reinterpret_cast<float*>(w);
  w += update_ws_kappa_bytes(a.B);
  a.counters = reinterpret_cast<int*>(w);
  a.nblk = update_blocks_per_sample(a.D, kSegSteps);  // a function of D only, never of B (shard-invariant reduction tree)
  dim3 grid(a.nblk, a.B);
  if (a.mode == 1) {  // AND: Gram pass + per-sample solve write kappa_in before the update pass reads it
    switch (a.M) {
      case 1: superpose_and_gram_kernel<1><<<grid, kUpdThreads, 0, st>>>(a); superpose_and_solve_kernel<1><<<a.B, 256, 0, st>>>(a); break;
      case 2: superpose_and_gram_kernel<2><<<grid, kUpdThreads, 0, st>>>(a); superpose_and_solve_kernel<2><<<a.B, 256, 0, st>>>(a); break;
      case 3: superpose_and_gram_kernel<3><<<grid, kUpdThreads, 0, st>>>(a); superpose_and_solve_kernel<3><<<a.B, 256, 0, st>>>(a); break;
      default: superpose_and_gram_kernel<4><<<grid, kUpdThreads, 0, st>>>(a); superpose_and_solve_kernel<4><<<a.B, 256, 0, st>>>(a); break;
    }
    ++g_launches;
    SDD_LAUNCH_CHECK();
  }
  switch (a.M) {
    case 1: superpose_update_kernel<1, kSegSteps><<<grid, kUpdThreads, 0, st>>>(a); break;
    case 2: superpose_update_kernel<2, kSegSteps><<<grid, kUpdThreads, 0, st>>>(a); break;
    case 3: superpose_update_kernel<3, kSegSteps><<<grid, kUpdThreads, 0, st>>>(a); break;
    default: superpose_update_kernel<4, kSegSteps><<<grid, kUpdThreads, 0, st>>>(a); break;
  }
  SDD_LAUNCH_CHECK();
  return SDD_OK;
}

__global__ void philox_normal_kernel(float* out, int B, int D, uint64_t seed, int64_t sample_offset, int draw) {
  const int nq = D >> 2;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * nq) return;
  int b = (int)(i / nq), q = (int)(i % nq);
  float4 v = philox_normal4(seed, (uint32_t)q, (uint32_t)(sample_offset + b), (uint32_t)draw);
  reinterpret_cast<float4*>(out)[i] = v;
}
__global__ void copy_f32_kernel(float* dst, const float* src, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}
// start of a sampler run: this call's parameters into the device block the captured step graph reads, step = 0
__global__ void begin_run_kernel(RunParams* dst, RunParams v, int* step) {
  *dst = v;
  *step = 0;
}

static int update_op(const float* x_in, float* x_out, const float* eps, const float* noise, const float* logq,
                     float* logq_out, float* kappa_out, float* xstats_out, int B, int D, int M, float alpha,
                     float alpha_bar, float beta, float temperature, const float* bias, uint64_t seed,
                     int64_t sample_offset, int draw_index, int mode, void* workspace, size_t workspace_bytes,
                     void* stream) {
  SDD_CHECK(x_in && x_out && eps && logq && logq_out && workspace, "null argument");
  SDD_CHECK(B > 0 && D > 0 && M >= 1 && M <= kMaxModels, "bad shape");
  SDD_CHECK(workspace_bytes >= update_workspace_bytes(B, D, M), "workspace too small");
  SDD_TRY(device_check());
  UpdateArgs a;
  memset(&a, 0, sizeof(a));
  a.x_in = x_in; a.x_out = x_out; a.eps = eps;
  a.logq = logq; a.logq_out = logq_out; a.kappa_out = kappa_out; a.xstats_out = xstats_out;
  a.sc.alpha = alpha; a.sc.alpha_bar = alpha_bar; a.sc.beta = beta;
  a.sc.draw_index = noise ? 0 : draw_index;
  step_scalars_fill(a.sc);
  a.rv.noise = noise; a.rv.noise_step_stride = 0; a.rv.seed = seed; a.rv.sample_offset = sample_offset;
  a.rv.temperature = temperature; a.rv.bias = bias;
  a.B = B; a.D = D; a.M = M; a.mode = mode;
  return launch_superpose_update(a, workspace, (cudaStream_t)stream);
}
}  // namespace sdd

extern "C" {

int sdd_superpose_update(const float* x_in, float* x_out, const float* eps, const float* noise, const float* logq,
                         float* logq_out, float* kappa_out, float* xstats_out, int B, int D, int M, float alpha,
                         float alpha_bar, float beta, float temperature, const float* bias, uint64_t seed,
                         int64_t sample_offset, int draw_index, void* workspace, size_t workspace_bytes,
                         void* stream) {
  return update_op(x_in, x_out, eps, noise, logq, logq_out, kappa_out, xstats_out, B, D, M, alpha, alpha_bar, beta,
                   temperature, bias, seed, sample_offset, draw_index, 0, workspace, workspace_bytes, stream);
}

int sdd_superpose_update_and(const float* x_in, float* x_out, const float* eps, const float* noise, const float* logq,
                             float* logq_out, float* kappa_out, float* xstats_out, int B, int D, int M, float alpha,
                             float alpha_bar, float beta, uint64_t seed, int64_t sample_offset, int draw_index,
                             void* workspace, size_t workspace_bytes, void* stream) {
  return update_op(x_in, x_out, eps, noise, logq, logq_out, kappa_out, xstats_out, B, D, M, alpha, alpha_bar, beta, 1.0f,
                   nullptr, seed, sample_offset, draw_index, 1, workspace, workspace_bytes, stream);
}

int sdd_attention_fwd(const void* q, const void* k, const void* vt, void* out, int BH, int S, int head_dim,
                      float scale, void* stream) {
  SDD_CHECK(q && k && vt && out, "null argument");
  SDD_CHECK(head_dim == kAttnD, "head_dim must be 64");
  SDD_CHECK(BH > 0 && S >= kAttnBN && S % kAttnBN == 0, "S must be a positive multiple of 128");
  SDD_CHECK(BH <= 65535, "batch * heads must be <= 65535");
  SDD_TRY(device_check());
  CUtensorMap tmQ, tmK, tmVt;
  SDD_TRY(make_attn_map(&tmQ, q, BH, S, kAttnD, kAttnBM));
  SDD_TRY(make_attn_map(&tmK, k, BH, S, kAttnD, kAttnBN));
  SDD_TRY(make_attn_map(&tmVt, vt, BH, kAttnD, S, kAttnD));
  SDD_TRY(ensure_func_attrs());
  AttnArgs a;
  a.out = reinterpret_cast<act_t*>(out); a.S = S; a.BH = BH;
  a.scale_log2e = scale * 1.4426950408889634f;
  attention_fwd_kernel<<<dim3((unsigned)(S / kAttnBM), (unsigned)BH), kAttnThreads, kAttnSmem, (cudaStream_t)stream>>>(
      tmQ, tmK, tmVt, a);
  SDD_LAUNCH_CHECK();
  return SDD_OK;
}

int sdd_attention_block_nhwc(const void* x, const float* gn_gamma, const float* gn_beta, const float* w_qkv,
                             const float* b_qkv, const float* w_out, const float* b_out, void* out, int B, int S, int C,
                             int heads, void* stream) {
  SDD_CHECK(x && gn_gamma && gn_beta && w_qkv && b_qkv && w_out && b_out && out, "null argument");
  SDD_CHECK(C == kAbC && heads == kAbHeads, "attention block supports C = 128 (2 heads of 64)");
  SDD_CHECK(B > 0 && S >= kAttnBN && S % kAttnBN == 0, "S must be a positive multiple of 128");
  SDD_TRY(device_check());
  cudaStream_t st = (cudaStream_t)stream;
  // operator form: scratch is allocated and freed here (the UNet variant passes its own arena and never synchronises)
  const size_t per = (size_t)B * S * kAbC;  // elements of one [B, S, 128] fp16 tensor
  act_t* buf = nullptr; float* mr = nullptr;
  if (cudaMalloc(&buf, 4 * per * sizeof(act_t)) != cudaSuccess || cudaMalloc(&mr, (size_t)B * 8 * sizeof(float)) != cudaSuccess) {
    cudaFree(buf); cudaFree(mr);
    set_error("cudaMalloc failed"); return SDD_ENOMEM;
  }
  gn_stats_nhwc_kernel<<<B * 4, 256, 0, st>>>((const act_t*)x, mr, S, kAbC);
  ++g_launches;
  int rc = launch_attention_block((const act_t*)x, nullptr, mr, AttnParams{gn_gamma, gn_beta, w_qkv, b_qkv, w_out, b_out},
                                  buf, buf + per, buf + 2 * per, buf + 3 * per, (act_t*)out, nullptr, B, S, st);
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(buf); cudaFree(mr);
  if (rc != SDD_OK) return rc;
  if (e != cudaSuccess) { set_error(std::string("attention block: ") + cudaGetErrorString(e)); return SDD_ECUDA; }
  return SDD_OK;
}

int sdd_attention_profile(const void* q, const void* k, const void* vt, void* out, int BH, int S, int head_dim,
                          float scale, int iters, float* ms_host, void* stream) {
  SDD_CHECK(ms_host && iters > 0, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int rc = SDD_OK;
  for (int i = 0; rc == SDD_OK && i < 3; ++i) rc = sdd_attention_fwd(q, k, vt, out, BH, S, head_dim, scale, stream);
  cudaEventRecord(e0, st);
  for (int i = 0; rc == SDD_OK && i < iters; ++i) rc = sdd_attention_fwd(q, k, vt, out, BH, S, head_dim, scale, stream);
  cudaEventRecord(e1, st);
  if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("attention profile: kernel failed"); rc = SDD_ECUDA; }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (rc == SDD_OK) *ms_host = ms / iters;
  return rc;
}

int sdd_q_sample(const float* x_start, const float* noise, const float* sqrt_ab, const float* sqrt_1mab, float* out,
                 int B, int D, void* stream) {
  SDD_CHECK(x_start && noise && sqrt_ab && sqrt_1mab && out, "null argument");
  SDD_CHECK(B > 0 && D > 0 && D % 4 == 0, "D must be a positive multiple of 4");
  SDD_TRY(device_check());
  const int nq = D / 4;
  dim3 grid((unsigned)std::min((nq + 255) / 256, 4 * num_sms()), (unsigned)B);
  q_sample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_start, noise, sqrt_ab, sqrt_1mab, out, D);
  SDD_LAUNCH_CHECK();
  return SDD_OK;
}

size_t sdd_mse_workspace(void) { return kMseBlocks * sizeof(double); }

int sdd_mse(const float* pred, const float* target, size_t n, float* out, void* workspace, size_t workspace_bytes,
            void* stream) {
  SDD_CHECK(pred && target && out && workspace, "null argument");
  SDD_CHECK(n > 0 && n % 4 == 0, "n must be a positive multiple of 4");
  SDD_CHECK(workspace_bytes >= sdd_mse_workspace(), "workspace too small");
  SDD_TRY(device_check());
  const size_t n4 = n / 4;
  const int blocks = (int)std::min<size_t>((n4 + 255) / 256, (size_t)kMseBlocks);
  mse_partial_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pred, target, n4, reinterpret_cast<double*>(workspace));
  SDD_LAUNCH_CHECK();
  mse_final_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double*>(workspace), blocks, (double)n, out);
  SDD_LAUNCH_CHECK();
  return SDD_OK;
}

int sdd_philox_normal(float* out, int B, int D, uint64_t seed, int64_t sample_offset, int draw_index, void* stream) {
  SDD_CHECK(out && B > 0 && D > 0 && D % 4 == 0, "bad argument");
  SDD_TRY(device_check());
  size_t n = (size_t)B * (D / 4);
  philox_normal_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out, B, D, seed, sample_offset,
                                                                                     draw_index);
  SDD_LAUNCH_CHECK();
  return SDD_OK;
}

}  // extern "C"

// =============================================================================== sampler
struct sdd_sampler {
  int M = 0, T = 0, B = 0, H = 0, W = 0, D = 0;
  sdd_unet* models[kMaxModels] = {nullptr, nullptr, nullptr, nullptr};
  float* tables[kMaxModels] = {nullptr, nullptr, nullptr, nullptr};  // [T][385], row = loop iteration
  StepScalars* sched = nullptr;  // [T], row = loop iteration (t = T-1-row)
  int* step = nullptr;
  RunParams* rp = nullptr;       // per-call parameters, rewritten by begin_run_kernel
  float *x = nullptr, *eps = nullptr, *logq = nullptr, *xstats = nullptr;  // logq: [2][B][M], double-buffered by step parity
  void* upd_ws = nullptr;
  float* stat_partials = nullptr; int* stat_counters = nullptr;
  cudaStream_t work = nullptr;   // private stream: graph capture is illegal on the legacy default stream
  cudaStream_t side[kMaxModels] = {nullptr, nullptr, nullptr, nullptr};  // models 1.. run beside model 0 (enqueue_step)
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxModels] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  cudaGraphExec_t exec = nullptr;
  // streamed host-noise path (sdd_sample_args::noise_host): a ring of two chunks of noise slices refilled on `copy`
  float* ring = nullptr; int ring_cap = 0;  // slices allocated
  cudaStream_t copy = nullptr;
  cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  int captured_mode = -1;        // the only per-call argument that changes the graph's topology
  int64_t captured_gen[kMaxModels] = {0, 0, 0, 0};
  int64_t launches_per_step = 0, launches_fixed = 0, graph_instantiations = 0;
};

namespace {

// One sampling step.  The M forwards are independent (they read x, write their own eps-hat, use their own workspaces), so
// models 1.. are enqueued on side streams forked from / joined into the work stream (captured as parallel branches of the
// step graph).  Every conv launch is a persistent grid with one CTA per SM, so the branches cannot share an SM -- what
// the fork buys is that a kernel's tail (SMs idle while the last tiles finish) and the next kernel's prologue (resident
// weights, TMEM allocation) are filled with the other model's CTAs instead of a drain between dependent launches, and
// the tiny launches (scale / shift tables, memsets, 1->1 conv) hide under the other branch's convs.
// A handle that appears twice (self-superposition) owns ONE workspace: then everything stays on the work stream.
int enqueue_step(sdd_sampler* s, int mode, bool first_step, cudaStream_t st) {
  bool fork = s->M > 1;
  for (int m = 0; m < s->M; ++m)
    for (int j = 0; j < m; ++j) fork = fork && s->models[m] != s->models[j];
  // GroupNorm(1,1) statistics of x: step 0 -> stats_x_kernel's (mean, rstd) of x_T; later steps -> the partial sums the
  // previous update launch left behind (deferred finalisation, update.cuh)
  XStatsSrc xs;
  memset(&xs, 0, sizeof(xs));
  if (first_step) {
    xs.xstats = s->xstats;
  } else {
    xs.partials = reinterpret_cast<const float*>(s->upd_ws);
    xs.parity_stride = update_ws_part_bytes(s->B, s->D) / sizeof(float);
    xs.nblk = update_blocks_per_sample(s->D, kSegSteps); xs.per_block = kPartialsPerBlock; xs.off = 3 * s->M;
    xs.step_ptr = s->step; xs.count = (float)s->D;
  }
  if (fork) SDD_CUDA(cudaEventRecord(s->ev_fork, st));
  for (int m = s->M - 1; m >= 0; --m) {
    cudaStream_t ms = (m == 0 || !fork) ? st : s->side[m];
    if (ms != st) SDD_CUDA(cudaStreamWaitEvent(ms, s->ev_fork, 0));
    BiasRef tb{s->tables[m], s->step, s->models[m]->bias_row, 0};
    SDD_TRY(unet_forward_impl(s->models[m], s->x, xs, tb, s->eps + (size_t)m * s->B * s->D, s->B, s->H, s->W, ms));
    if (ms != st) SDD_CUDA(cudaEventRecord(s->ev_join[m], ms));
  }
  for (int m = 1; fork && m < s->M; ++m) SDD_CUDA(cudaStreamWaitEvent(st, s->ev_join[m], 0));
  UpdateArgs a;
  memset(&a, 0, sizeof(a));
  a.x_in = s->x; a.x_out = s->x; a.eps = s->eps;
  a.logq = s->logq; a.logq_out = s->logq; a.kappa_out = nullptr; a.xstats_out = nullptr;
  a.table = s->sched; a.step_ptr = s->step; a.advance_step = 1; a.defer = 1;
  a.rp = s->rp;
  a.B = s->B; a.D = s->D; a.M = s->M; a.mode = mode;
  return launch_superpose_update(a, s->upd_ws, st);
}

}  // namespace

extern "C" {

int sdd_sampler_create(sdd_sampler_t** out, sdd_unet_t* const* models, int M, const float* alphas_host,
                       const float* alpha_bars_host, const float* betas_host, int T, int B, int H, int W,
                       void* stream) {
  SDD_CHECK(out && models && alphas_host && alpha_bars_host && betas_host, "null argument");
  SDD_CHECK(M >= 1 && M <= kMaxModels, "1 <= M <= 4");
  SDD_CHECK(T >= 1 && B >= 1, "bad T or B");
  SDD_CHECK(H % kTileH == 0 && W % kTileW == 0 && H >= 16 && W >= 8, "H must be a multiple of 16 and W of 8");
  SDD_TRY(device_check());
  cudaStream_t st = (cudaStream_t)stream;
  sdd_sampler* s = new sdd_sampler();
  s->M = M; s->T = T; s->B = B; s->H = H; s->W = W; s->D = H * W;
  auto fail = [&](int code) { sdd_sampler_destroy(s); return code; };
#define S_CUDA(expr) do { if ((expr) != cudaSuccess) { set_error(#expr " failed"); return fail(SDD_ECUDA); } } while (0)
  const size_t BD = (size_t)B * s->D;
  S_CUDA(cudaStreamCreateWithFlags(&s->work, cudaStreamNonBlocking));
  S_CUDA(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
  for (int m = 1; m < M; ++m) {
    S_CUDA(cudaStreamCreateWithFlags(&s->side[m], cudaStreamNonBlocking));
    S_CUDA(cudaEventCreateWithFlags(&s->ev_join[m], cudaEventDisableTiming));
  }
  S_CUDA(cudaEventCreateWithFlags(&s->ev_in, cudaEventDisableTiming));
  S_CUDA(cudaEventCreateWithFlags(&s->ev_out, cudaEventDisableTiming));
  S_CUDA(cudaMalloc(&s->x, BD * sizeof(float)));
  S_CUDA(cudaMalloc(&s->eps, (size_t)M * BD * sizeof(float)));
  S_CUDA(cudaMalloc(&s->logq, (size_t)2 * B * M * sizeof(float)));
  S_CUDA(cudaMalloc(&s->xstats, (size_t)B * 2 * sizeof(float)));
  S_CUDA(cudaMalloc(&s->step, sizeof(int)));
  S_CUDA(cudaMalloc(&s->rp, sizeof(RunParams)));
  size_t uw = update_workspace_bytes(B, s->D, M);
  S_CUDA(cudaMalloc(&s->upd_ws, uw));
  S_CUDA(cudaMemsetAsync(s->upd_ws, 0, uw, st));
  S_CUDA(cudaMalloc(&s->stat_partials, (size_t)B * kStatsBlocks * 2 * sizeof(float)));
  S_CUDA(cudaMalloc(&s->stat_counters, (size_t)B * sizeof(int)));
  S_CUDA(cudaMemsetAsync(s->stat_counters, 0, (size_t)B * sizeof(int), st));
  // schedule rows indexed by loop iteration: row k <-> t = T-1-k (ddpm.py:34)
  std::vector<StepScalars> sc(T);
  std::vector<int64_t> trev(T);
  for (int k = 0; k < T; ++k) {
    int t = T - 1 - k;
    sc[k].alpha = alphas_host[t]; sc[k].alpha_bar = alpha_bars_host[t]; sc[k].beta = betas_host[t];
    sc[k].draw_index = t > 0 ? k + 1 : -1;  // ddpm.py:36: no noise at t == 0
    step_scalars_fill(sc[k]);
    trev[k] = t;
  }
  S_CUDA(cudaMalloc(&s->sched, T * sizeof(StepScalars)));
  S_CUDA(cudaMemcpyAsync(s->sched, sc.data(), T * sizeof(StepScalars), cudaMemcpyHostToDevice, st));
  int64_t* t_dev = nullptr;
  float *emb0 = nullptr, *h1 = nullptr, *emb = nullptr;
  S_CUDA(cudaMalloc(&t_dev, T * sizeof(int64_t)));
  S_CUDA(cudaMemcpyAsync(t_dev, trev.data(), T * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  S_CUDA(cudaMalloc(&emb0, (size_t)T * kTimeDim * sizeof(float)));
  S_CUDA(cudaMalloc(&h1, (size_t)T * 4 * kTimeDim * sizeof(float)));
  S_CUDA(cudaMalloc(&emb, (size_t)T * kTimeDim * sizeof(float)));
  int rc = SDD_OK;
  for (int m = 0; m < M && rc == SDD_OK; ++m) {
    s->models[m] = models[m];
    if (!models[m]) { set_error("null model"); rc = SDD_EINVAL; break; }
    if (cudaMalloc(&s->tables[m], (size_t)T * models[m]->bias_row * sizeof(float)) != cudaSuccess) { set_error("cudaMalloc(table) failed"); rc = SDD_ENOMEM; break; }
    // every timestep's (conv2 bias + time embedding) rows, once: t is batch-uniform in sampling (ddpm.py:35)
    rc = time_bias_rows(models[m], t_dev, nullptr, T, emb0, h1, emb, s->tables[m], st);
    if (rc == SDD_OK) rc = ensure_workspace(models[m], B, H, W);
  }
  cudaError_t se = cudaStreamSynchronize(st);
  cudaFree(t_dev); cudaFree(emb0); cudaFree(h1); cudaFree(emb);
  if (rc != SDD_OK) return fail(rc);
  if (se != cudaSuccess) { set_error(std::string("sampler_create sync: ") + cudaGetErrorString(se)); return fail(SDD_ECUDA); }
#undef S_CUDA
  *out = s;
  return SDD_OK;
}

int sdd_sampler_run(sdd_sampler_t* s, const sdd_sample_args* args, void* stream) {
  SDD_CHECK(s && args && args->x_out, "null argument");
  SDD_CHECK(args->mode == 0 || args->mode == 1, "mode must be 0 (OR) or 1 (AND)");
  cudaStream_t user = (cudaStream_t)stream;
  cudaStream_t st = s->work;  // everything runs here, ordered after / before the caller's stream by events
  SDD_CUDA(cudaEventRecord(s->ev_in, user));
  SDD_CUDA(cudaStreamWaitEvent(st, s->ev_in, 0));
  const size_t BD = (size_t)s->B * s->D;
  const int64_t l0 = g_launches;
  // --- this call's parameters -> device block read by the (possibly already instantiated) step graph; step = 0
  // --- streamed host noise: slice j of the pinned host stack lives in ring slot j % (2 S); chunk c = slices [c S, (c+1) S)
  // is copied on `copy` into ring half c & 1 once the steps that read chunk c-2 have finished (ev_free), and the step
  // that needs its first slice waits for it (ev_ready).  All of it is enqueued up front: no host synchronisation.
  const float* noise_dev = args->noise_stack;
  int S = 0, nchunks = 0, ring_n = 0;  // slices per chunk, chunks, slots the kernels index modulo
  SDD_CHECK(!(args->noise_stack && args->noise_host), "pass noise_stack or noise_host, not both");
  if (args->noise_host) {
    cudaPointerAttributes pa;
    SDD_CHECK(cudaPointerGetAttributes(&pa, args->noise_host) == cudaSuccess && pa.type == cudaMemoryTypeHost,
              "noise_host must be page-locked (pinned) host memory");
    const size_t slice_bytes = BD * sizeof(float);
    S = args->noise_host_chunk > 0 ? args->noise_host_chunk : (int)std::max<size_t>(1, ((size_t)64 << 20) / slice_bytes);
    S = std::min(S, s->T);
    ring_n = 2 * S >= s->T ? s->T : 2 * S;  // (the whole stack fits in two chunks: slot j = j)
    if (s->ring_cap < ring_n) {
      SDD_CUDA(cudaStreamSynchronize(s->work));
      cudaFree(s->ring); s->ring = nullptr; s->ring_cap = 0;
      if (cudaMalloc(&s->ring, (size_t)ring_n * slice_bytes) != cudaSuccess) { set_error("cudaMalloc(noise ring) failed"); return SDD_ENOMEM; }
      s->ring_cap = ring_n;
    }
    if (!s->copy) {
      SDD_CUDA(cudaStreamCreateWithFlags(&s->copy, cudaStreamNonBlocking));
      for (int i = 0; i < 2; ++i) {
        SDD_CUDA(cudaEventCreateWithFlags(&s->ev_ready[i], cudaEventDisableTiming));
        SDD_CUDA(cudaEventCreateWithFlags(&s->ev_free[i], cudaEventDisableTiming));
      }
    }
    nchunks = (s->T + S - 1) / S;
    noise_dev = s->ring;
    // the ring's previous readers: the last run on this sampler (ev_out: recorded on the work stream at its end; a no-op
    // before the first run) and whatever the caller ordered before this call (ev_in)
    SDD_CUDA(cudaStreamWaitEvent(s->copy, s->ev_out, 0));
    SDD_CUDA(cudaStreamWaitEvent(s->copy, s->ev_in, 0));
  }
  auto enqueue_chunk = [&](int c) -> int {  // H2D copy of chunk c into its ring half
    if (c >= nchunks) return SDD_OK;
    if (c >= 2) SDD_CUDA(cudaStreamWaitEvent(s->copy, s->ev_free[c & 1], 0));
    const int j0 = c * S, n = std::min(S, s->T - j0);
    const int slot0 = ring_n == s->T ? j0 : (c & 1) * S;
    SDD_CUDA(cudaMemcpyAsync(s->ring + (size_t)slot0 * BD, args->noise_host + (size_t)j0 * BD, (size_t)n * BD * sizeof(float),
                             cudaMemcpyHostToDevice, s->copy));
    SDD_CUDA(cudaEventRecord(s->ev_ready[c & 1], s->copy));
    return SDD_OK;
  };
  // before the step that reads slice j: its chunk must have landed; the NEXT chunk's copy is enqueued at the same time,
  // so that it runs under this chunk's steps.  after_slice(j): the last reader of a chunk releases its ring half.
  auto before_slice = [&](int j) -> int {
    if (!args->noise_host || j >= s->T || j % S != 0) return SDD_OK;
    const int c = j / S;
    if (c == 0) { SDD_TRY(enqueue_chunk(0)); }
    SDD_TRY(enqueue_chunk(c + 1));
    SDD_CUDA(cudaStreamWaitEvent(st, s->ev_ready[c & 1], 0));
    return SDD_OK;
  };
  auto after_slice = [&](int j) -> int {
    if (!args->noise_host || j >= s->T) return SDD_OK;
    if (j % S == S - 1 || j == s->T - 1) SDD_CUDA(cudaEventRecord(s->ev_free[(j / S) & 1], st));
    return SDD_OK;
  };
  RunParams rv;
  memset(&rv, 0, sizeof(rv));
  rv.noise = noise_dev; rv.noise_step_stride = (int64_t)BD;
  rv.noise_ring = ring_n;
  rv.seed = args->seed; rv.sample_offset = args->sample_offset; rv.temperature = args->temperature;
  rv.bias = args->bias; rv.kappa_traj = args->kappa_traj; rv.logq_traj = args->logq_traj; rv.x_traj = args->x_traj;
  begin_run_kernel<<<1, 1, 0, st>>>(s->rp, rv, s->step);
  SDD_LAUNCH_CHECK();
  // --- x_T, logq = 0, GN(1,1) stats of x_T
  SDD_CUDA(cudaMemsetAsync(s->logq, 0, (size_t)2 * s->B * s->M * sizeof(float), st));
  if (args->logq_traj) SDD_CUDA(cudaMemsetAsync(args->logq_traj, 0, (size_t)s->B * s->M * sizeof(float), st));
  if (noise_dev) {
    SDD_TRY(before_slice(0));  // slice 0 = x_T
    SDD_CUDA(cudaMemcpyAsync(s->x, noise_dev, BD * sizeof(float), cudaMemcpyDeviceToDevice, st));
    SDD_TRY(after_slice(0));
  } else {
    size_t n = BD / 4;
    philox_normal_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s->x, s->B, s->D, args->seed,
                                                                      args->sample_offset, 0);
    SDD_LAUNCH_CHECK();
  }
  stats_x_kernel<<<dim3(kStatsBlocks, s->B), 256, 0, st>>>(s->x, s->D, s->stat_partials, s->stat_counters, s->xstats);
  SDD_LAUNCH_CHECK();
  if (args->x_traj) SDD_CUDA(cudaMemcpyAsync(args->x_traj, s->x, BD * sizeof(float), cudaMemcpyDeviceToDevice, st));
  const int64_t l1 = g_launches;
  for (int m = 0; m < s->M; ++m) SDD_TRY(ensure_workspace(s->models[m], s->B, s->H, s->W));
  // --- T steps.  Step 0 always runs eagerly (so every kernel is loaded before a capture begins).
  {
    const int64_t lk = g_launches;
    SDD_TRY(before_slice(1));  // step k reads slice k + 1 (the last step, t == 0, reads none)
    SDD_TRY(enqueue_step(s, args->mode, true, st));
    SDD_TRY(after_slice(1));
    s->launches_per_step = g_launches - lk;
  }
  if (args->use_graph && s->T > 1) {
    // The graph holds only pointers to sampler-owned device state (x, eps, logq, step counter, RunParams block, bias
    // tables): seed, shard offset, noise stack, temperature, bias and trajectory buffers change WITHOUT a re-capture.
    bool stale = !s->exec || s->captured_mode != args->mode;
    for (int m = 0; m < s->M; ++m) stale = stale || s->captured_gen[m] != ws_generation(s->models[m]);
    if (stale) {
      if (s->exec) { cudaGraphExecDestroy(s->exec); s->exec = nullptr; }
      cudaGraph_t graph = nullptr;
      const int64_t lk = g_launches;
      SDD_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      int rc = enqueue_step(s, args->mode, false, st);
      cudaError_t ce = cudaStreamEndCapture(st, &graph);  // always leave capture mode, even on error
      g_launches = lk;                                    // captured, not launched
      if (rc != SDD_OK) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc; }
      SDD_CUDA(ce);
      ce = cudaGraphInstantiate(&s->exec, graph, 0);
      cudaGraphDestroy(graph);
      SDD_CUDA(ce);
      s->captured_mode = args->mode;
      ++s->graph_instantiations;
      for (int m = 0; m < s->M; ++m) s->captured_gen[m] = ws_generation(s->models[m]);
    }
    for (int k = 1; k < s->T; ++k) {
      SDD_TRY(before_slice(k + 1));
      SDD_CUDA(cudaGraphLaunch(s->exec, st));
      SDD_TRY(after_slice(k + 1));
    }
  } else {
    for (int k = 1; k < s->T; ++k) {
      SDD_TRY(before_slice(k + 1));
      SDD_TRY(enqueue_step(s, args->mode, false, st));
      SDD_TRY(after_slice(k + 1));
    }
  }
  {  // close the run: log q_T from the last step's partials (the per-step finalisation is deferred, update.cuh)
    UpdateArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.logq = s->logq; fa.logq_out = s->logq; fa.table = s->sched; fa.rp = s->rp;
    fa.partials = reinterpret_cast<float*>(s->upd_ws);
    fa.parity_stride = update_ws_part_bytes(s->B, s->D) / sizeof(float);
    fa.nblk = update_blocks_per_sample(s->D, kSegSteps);
    fa.B = s->B; fa.D = s->D; fa.M = s->M;
    switch (s->M) {
      case 1: finish_run_kernel<1><<<s->B, 32, 0, st>>>(fa, s->T); break;
      case 2: finish_run_kernel<2><<<s->B, 32, 0, st>>>(fa, s->T); break;
      case 3: finish_run_kernel<3><<<s->B, 32, 0, st>>>(fa, s->T); break;
      default: finish_run_kernel<4><<<s->B, 32, 0, st>>>(fa, s->T); break;
    }
    SDD_LAUNCH_CHECK();
  }
  s->launches_fixed = (l1 - l0) + 2;
  int blocks = (int)std::min<size_t>((BD + 255) / 256, 2048);
  copy_f32_kernel<<<blocks, 256, 0, st>>>(args->x_out, s->x, BD);
  SDD_LAUNCH_CHECK();
  SDD_CUDA(cudaEventRecord(s->ev_out, st));
  SDD_CUDA(cudaStreamWaitEvent(user, s->ev_out, 0));
  return SDD_OK;
}

int64_t sdd_sampler_launches_per_run(const sdd_sampler_t* s) {
  return s ? s->launches_fixed + (int64_t)s->T * s->launches_per_step : 0;
}

int64_t sdd_sampler_graph_instantiations(const sdd_sampler_t* s) { return s ? s->graph_instantiations : 0; }

int sdd_sampler_destroy(sdd_sampler_t* s) {
  if (!s) return SDD_OK;
  if (s->work) cudaStreamSynchronize(s->work);
  for (int m = 1; m < kMaxModels; ++m) {
    if (s->side[m]) { cudaStreamSynchronize(s->side[m]); cudaStreamDestroy(s->side[m]); }
    if (s->ev_join[m]) cudaEventDestroy(s->ev_join[m]);
  }
  if (s->ev_fork) cudaEventDestroy(s->ev_fork);
  if (s->copy) { cudaStreamSynchronize(s->copy); cudaStreamDestroy(s->copy); }
  for (int i = 0; i < 2; ++i) {
    if (s->ev_ready[i]) cudaEventDestroy(s->ev_ready[i]);
    if (s->ev_free[i]) cudaEventDestroy(s->ev_free[i]);
  }
  cudaFree(s->ring);
  if (s->exec) cudaGraphExecDestroy(s->exec);
  if (s->ev_in) cudaEventDestroy(s->ev_in);
  if (s->ev_out) cudaEventDestroy(s->ev_out);
  if (s->work) cudaStreamDestroy(s->work);
  for (int m = 0; m < kMaxModels; ++m) cudaFree(s->tables[m]);
  cudaFree(s->sched); cudaFree(s->step); cudaFree(s->rp); cudaFree(s->x); cudaFree(s->eps); cudaFree(s->logq);
  cudaFree(s->xstats); cudaFree(s->upd_ws); cudaFree(s->stat_partials); cudaFree(s->stat_counters);
  delete s;
  return SDD_OK;
}

// ------------------------------------------------------------------------------- operator entry points
#ifdef SDD_CONV_PROF
// prof build only (tools/conv_prof.py): read and clear the conv kernel's per-role cycle counters
extern "C" int sdd_conv_prof_read(unsigned long long* out64) {
  SDD_CUDA(cudaDeviceSynchronize());
  SDD_CUDA(cudaMemcpyFromSymbol(out64, g_conv_prof, sizeof(unsigned long long) * 64));
  unsigned long long z[64] = {};
  SDD_CUDA(cudaMemcpyToSymbol(g_conv_prof, z, sizeof(z)));
  return SDD_OK;
}
#endif

// Kernel-only timing for the roofline: `iters` launches of the product conv at one shape, each bracketed by CUDA events
// on the launching stream, with `flush_bytes` of `flush` rewritten before every launch (L2 flush).
int sdd_conv3x3_profile(const void* act, const float* w, const float* bias, void* out, int B, int H, int W, int Cin,
                        int Cout, int impl, int iters, void* flush, size_t flush_bytes, float* ms_host, void* stream) {
  SDD_CHECK(act && w && bias && out && ms_host && iters > 0, "bad argument");
  SDD_CHECK(impl == 1 || impl == 2, "impl: 1 = conv alone, 2 = with the fused GroupNorm+SiLU");
  SDD_CHECK((Cin == 64 || Cin == 128) && (Cout == 64 || Cout == 128), "Cin, Cout must be 64 or 128");
  SDD_CHECK(H % kTileH == 0 && W % kTileW == 0, "H must be a multiple of 16 and W a multiple of 8");
  SDD_TRY(device_check());
  cudaStream_t st = (cudaStream_t)stream;
  act_t* wt = nullptr;
  float* gin = nullptr;  // identity GroupNorm input: mean 0, rstd 1 | gamma 1 | beta 0
  long long* osums = nullptr;
  float* gn_ab = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const int total_w = 9 * Cout * Cin;
  int rc = SDD_OK;
  CUtensorMap tmA, tmB;
  if (cudaMalloc(&wt, (size_t)total_w * 2) != cudaSuccess || cudaMalloc(&gin, ((size_t)B * 8 + 256) * 4) != cudaSuccess ||
      cudaMalloc(&osums, (size_t)B * 8 * sizeof(long long)) != cudaSuccess ||
      cudaMalloc(&gn_ab, (size_t)2 * B * Cin * sizeof(float)) != cudaSuccess) {
    set_error("cudaMalloc failed"); rc = SDD_ENOMEM;
  }
  if (rc == SDD_OK) {
    std::vector<float> h((size_t)B * 8 + 256, 0.f);
    for (int i = 0; i < B * 4; ++i) h[2 * i + 1] = 1.f;
    for (int i = 0; i < 128; ++i) h[(size_t)B * 8 + i] = 1.f;
    cudaMemcpyAsync(gin, h.data(), h.size() * 4, cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    conv_weight_to_act_kernel<<<(total_w + 255) / 256, 256, 0, st>>>(w, wt, Cout, Cin);
    rc = make_act_map(&tmA, act, B, H, W, Cin);
  }
  if (rc == SDD_OK) rc = make_wt_map(&tmB, wt, Cout, Cin);
  GnInput3 gi{nullptr, nullptr, nullptr, nullptr, nullptr};
  if (impl == 2) gi = GnInput3{nullptr, gin, gin + (size_t)B * 8, gin + (size_t)B * 8 + 128, gn_ab};
  if (rc == SDD_OK && (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess)) rc = SDD_ECUDA;
  BiasRef br{bias, nullptr, 0, 0};
  double total = 0.0;
  for (int i = 0; rc == SDD_OK && i < iters + 2; ++i) {
    if (flush && flush_bytes) cudaMemsetAsync(flush, i & 0xff, flush_bytes, st);
    cudaMemsetAsync(osums, 0, (size_t)B * 8 * sizeof(long long), st);
    cudaEventRecord(e0, st);
    rc = launch_conv(tmA, tmB, (const act_t*)act, (act_t*)out, br, gi, osums, B, H, W, Cin, Cout, st);
    cudaEventRecord(e1, st);
    if (cudaEventSynchronize(e1) != cudaSuccess) { set_error("conv profile: kernel failed"); rc = SDD_ECUDA; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (i >= 2) total += ms;
  }
  cudaStreamSynchronize(st);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  cudaFree(osums); cudaFree(wt); cudaFree(gin); cudaFree(gn_ab);
  if (rc == SDD_OK) *ms_host = (float)(total / iters);
  return rc;
}

// The fused update step (ONE launch: HBM pass + per-sample finalize by the last-arriving CTA), M models, fp32 eps;
// noise == NULL => in-kernel Philox.  One event pair per launch, `flush` rewritten before each (also evicts the code).
int sdd_superpose_update_profile(float* x, const float* eps, const float* noise, float* logq, int B, int D, int M,
                                 int iters, void* flush, size_t flush_bytes, float* ms_host, void* stream) {
  SDD_CHECK(x && eps && logq && ms_host && iters > 0, "bad argument");
  SDD_TRY(device_check());
  cudaStream_t st = (cudaStream_t)stream;
  void* ws = nullptr;
  size_t wb = update_workspace_bytes(B, D, M);
  SDD_CUDA(cudaMalloc(&ws, wb));
  cudaMemsetAsync(ws, 0, wb, st);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  UpdateArgs a;
  memset(&a, 0, sizeof(a));
  a.x_in = x; a.x_out = x; a.eps = eps; a.logq = logq; a.logq_out = logq;
  a.rv.noise = noise; a.rv.temperature = 1.0f; a.rv.seed = 1234;
  a.sc.alpha = 0.99f; a.sc.alpha_bar = 0.5f; a.sc.beta = 0.01f;
  step_scalars_fill(a.sc);
  a.B = B; a.D = D; a.M = M;
  int rc = SDD_OK;
  double total = 0.0;
  for (int i = 0; rc == SDD_OK && i < iters + 2; ++i) {
    if (flush && flush_bytes) cudaMemsetAsync(flush, i & 0xff, flush_bytes, st);
    a.sc.draw_index = noise ? 0 : i;
    cudaEventRecord(e0, st);
    rc = launch_superpose_update(a, ws, st);
    cudaEventRecord(e1, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("update profile: kernel failed"); rc = SDD_ECUDA; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (i >= 2) total += ms;
  }
  cudaStreamSynchronize(st);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(ws);
  if (rc == SDD_OK) *ms_host = (float)(total / iters);
  return rc;
}

int sdd_superpose_update_profile_rotating(int B, int D, int M, int use_noise, int iters, size_t rot_bytes,
                                          float* ms_host, void* stream) {
  SDD_CHECK(ms_host && iters > 0 && B > 0 && D > 0 && D % 4 == 0 && M >= 1 && M <= kMaxModels, "bad argument");
  SDD_TRY(device_check());
  cudaStream_t st = (cudaStream_t)stream;
  const size_t BD = (size_t)B * D;
  const size_t set_floats = BD * (size_t)(1 + M + (use_noise ? 1 : 0));
  int nsets = (int)((rot_bytes + set_floats * 4 - 1) / (set_floats * 4));
  nsets = nsets < 2 ? 2 : (nsets > 512 ? 512 : nsets);
  // exactly the sampler's launch: deferred finalisation, schedule table + device step counter, double-buffered log q
  float* pool = nullptr; float* logq = nullptr; void* ws = nullptr; StepScalars* table = nullptr; int* step = nullptr;
  const size_t wb = update_workspace_bytes(B, D, M);
  const int rows = nsets + iters + 8;
  if (cudaMalloc(&pool, set_floats * 4 * nsets) != cudaSuccess || cudaMalloc(&logq, (size_t)2 * B * M * 4) != cudaSuccess ||
      cudaMalloc(&ws, wb) != cudaSuccess || cudaMalloc(&table, (size_t)rows * sizeof(StepScalars)) != cudaSuccess ||
      cudaMalloc(&step, sizeof(int)) != cudaSuccess) {
    cudaFree(pool); cudaFree(logq); cudaFree(ws); cudaFree(table); cudaFree(step);
    set_error("cudaMalloc failed"); return SDD_ENOMEM;
  }
  cudaMemsetAsync(ws, 0, wb, st);
  cudaMemsetAsync(logq, 0, (size_t)2 * B * M * 4, st);
  cudaMemsetAsync(step, 0, sizeof(int), st);
  {
    std::vector<StepScalars> h(rows);
    for (int i = 0; i < rows; ++i) {
      h[i].alpha = 0.99f; h[i].alpha_bar = 0.5f; h[i].beta = 0.01f; h[i].draw_index = use_noise ? 0 : i;
      step_scalars_fill(h[i]);
    }
    cudaMemcpyAsync(table, h.data(), (size_t)rows * sizeof(StepScalars), cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
  }
  {  // standard normals everywhere (values only matter for not being denormal / NaN)
    const size_t total = set_floats * nsets;
    const size_t nq = total / 4;
    philox_normal_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(pool, 1, (int)(nq * 4 > 0x7ffffff0 ? 0x7ffffff0 : nq * 4), 99, 0, 0);
    if (nq * 4 > 0x7ffffff0) cudaMemsetAsync(pool + 0x7ffffff0, 0, (total - 0x7ffffff0) * 4, st);
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  UpdateArgs a;
  memset(&a, 0, sizeof(a));
  a.logq = logq; a.logq_out = logq;
  a.table = table; a.step_ptr = step; a.advance_step = 1; a.defer = 1;
  a.rv.temperature = 1.0f; a.rv.seed = 1234;
  a.B = B; a.D = D; a.M = M;
  int rc = SDD_OK;
  auto run = [&](int n, int first) {
    for (int i = 0; rc == SDD_OK && i < n; ++i) {
      float* set = pool + (size_t)((first + i) % nsets) * set_floats;
      a.x_in = set; a.x_out = set; a.eps = set + BD; a.rv.noise = use_noise ? set + BD * (1 + M) : nullptr;
      rc = launch_superpose_update(a, ws, st);
    }
  };
  run(nsets, 0);  // warm-up: code, constants, TLBs; every set touched once
  cudaEventRecord(e0, st);
  run(iters, nsets);
  cudaEventRecord(e1, st);
  if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("update profile: kernel failed"); rc = SDD_ECUDA; }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(ws); cudaFree(logq); cudaFree(pool); cudaFree(table); cudaFree(step);
  if (rc == SDD_OK) *ms_host = ms / iters;
  return rc;
}

int sdd_conv3x3_fused_nhwc(const void* act_raw, const float* in_meanrstd, const float* in_gamma, const float* in_beta,
                           const float* w, const float* bias, int64_t bias_batch_stride, void* out,
                           float* gn_meanrstd, int B, int H, int W, int Cin, int Cout, void* stream) {
  SDD_CHECK(act_raw && w && bias && out, "null argument");
  SDD_CHECK((Cin == 64 || Cin == 128) && (Cout == 64 || Cout == 128), "Cin, Cout must be 64 or 128");
  SDD_CHECK(H % kTileH == 0 && W % kTileW == 0, "H must be a multiple of 16 and W a multiple of 8");
  SDD_CHECK(!in_meanrstd || (in_gamma && in_beta), "in_gamma / in_beta required with in_meanrstd");
  SDD_TRY(device_check());
  cudaStream_t st = (cudaStream_t)stream;
  act_t* wt = nullptr; long long* osums = nullptr;
  const int total_w = 9 * Cout * Cin;
  int rc = SDD_OK;
  CUtensorMap tmA, tmB;
  if (cudaMalloc(&wt, (size_t)total_w * 2) != cudaSuccess || cudaMalloc(&osums, (size_t)B * 8 * sizeof(long long)) != cudaSuccess) {
    set_error("cudaMalloc failed"); rc = SDD_ENOMEM;
  }
  if (rc == SDD_OK) {
    cudaMemsetAsync(osums, 0, (size_t)B * 8 * sizeof(long long), st);
    conv_weight_to_act_kernel<<<(total_w + 255) / 256, 256, 0, st>>>(w, wt, Cout, Cin);
    ++g_launches;
    rc = make_act_map(&tmA, act_raw, B, H, W, Cin);
  }
  if (rc == SDD_OK) rc = make_wt_map(&tmB, wt, Cout, Cin);
  float* gn_ab = nullptr;
  if (rc == SDD_OK && in_meanrstd && cudaMalloc(&gn_ab, (size_t)2 * B * Cin * sizeof(float)) != cudaSuccess) {
    set_error("cudaMalloc failed"); rc = SDD_ENOMEM;
  }
  if (rc == SDD_OK)
    rc = launch_conv(tmA, tmB, (const act_t*)act_raw, (act_t*)out, BiasRef{bias, nullptr, 0, bias_batch_stride},
                     GnInput3{nullptr, in_meanrstd, in_gamma, in_beta, gn_ab}, osums, B, H, W, Cin, Cout, st);
  if (rc == SDD_OK && gn_meanrstd) {
    gn_sums_to_meanrstd_kernel<<<(B * 4 + 127) / 128, 128, 0, st>>>(osums, gn_meanrstd, B * 4,
                                                                   (double)H * (double)W * (double)(Cout / 4), kGnEps);
    ++g_launches;
  }
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(wt); cudaFree(osums); cudaFree(gn_ab);
  if (rc != SDD_OK) return rc;
  if (e != cudaSuccess) { set_error(std::string("conv3x3_fused: ") + cudaGetErrorString(e)); return SDD_ECUDA; }
  return SDD_OK;
}

}  // extern "C"
