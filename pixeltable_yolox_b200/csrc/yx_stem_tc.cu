// Fused Focus + stem conv for sm_100a: space-to-depth, 3x3 conv, folded BN and SiLU in ONE kernel.
//   reference: Focus.forward (yolox/models/network_blocks.py:186-208) followed by its BaseConv
//   (network_blocks.py:27-52).  Focus(TL,BL,TR,BR) + 3x3 stride-1 pad-1 conv over 12 channels is the
//   same linear map as a 6x6 stride-2 pad-2 conv over the 3 image channels:
//       W6[o, c, 2u+py, 2v+px] = W[o, 3*(2*px+py) + c, u, v].
// The raw NCHW image (fp32 or uint8, 0..255) is read once and the NHWC 16-bit stem output written
// once; nothing else touches HBM (the unfused path writes and re-reads a 12-channel tensor and pays
// nine 32-byte-row TMA gathers per pixel).
//
// GEMM view: M = output pixels (tile = 8 rows x 16 cols), K = 6*3*6 = 108 (zero padded to 128),
// N = out_c.  TMA cannot build this operand directly (fp32/u8 source, channel-planar layout), so:
//   warp 9      TMA-loads each tile's 3 x 20 x 36 input patch (one 4-D box of the NCHW image, zero
//               filled outside the image = Conv2d padding) into a 4-deep ring, keeping ~35 KB of
//               loads in flight per SM;
//   warps 0-7   two groups of "builders" (alternating tiles): convert the patch to bf16/fp16 and write the [128 x 128] K-major,
//               128B-swizzled A operand with 16-byte shared stores (fence.proxy.async publishes them
//               to the tensor core);
//   warp 8      issues tcgen05.mma against the weights (out_c x 128) resident in shared memory,
//               accumulators in TMEM (4 stages);
//   warps 10-17 two epilogue groups (bias + SiLU + 256-bit stores).
#include <stdlib.h>
#include <string.h>

#include "yx_tc_epilogue.cuh"

namespace yx {

static constexpr int kStemStages = 3;
static constexpr int kStemAcc = 4;
static constexpr int kStemBuilders = 128;            // threads per builder group (one operand row each)
static constexpr int kStemBuilderGroups = 2;         // tiles alternate between the groups
static constexpr int kStemEpiGroups = 2;
static constexpr int kStemThreads = kStemBuilders * kStemBuilderGroups + 64 + 128 * kStemEpiGroups;   // + MMA warp + TMA warp
static constexpr int kWarpMma = kStemBuilders * kStemBuilderGroups / 32, kWarpTma = kWarpMma + 1, kWarpEpi = kWarpMma + 2;
static constexpr int kPatchRows = 20;
static constexpr int kPatchStages = 4;
static constexpr int kPatchStageBytes = 9600;      // 3*20*40 fp32 (u8: 3*20*64 = 3840), 128-byte multiple
static constexpr int kTileH = 8, kTileW = 16;

struct StemParams {
  const void* img; int img_dtype;
  int batch, h, w;            // image size
  int out_h, out_w;           // h/2, w/2
  int tiles_w, tiles_h, num_tiles;
  int BN, BNpad;              // out_c (multiple of 16) and its TMEM pitch
  unsigned idesc, desc_hi, tmem_cols;
  unsigned bias_bytes, b_bytes, patch_tx;
  int debug;
  EpiParams epi;
};

struct __align__(8) StemShared {
  uint64_t full[kStemStages];
  uint64_t empty[kStemStages];
  uint64_t tmem_full[kStemAcc];
  uint64_t tmem_empty[kStemAcc];
  uint64_t w_full;
  uint64_t patch_full[kPatchStages];
  uint64_t patch_empty[kPatchStages];
  uint32_t tmem_base;
};

__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <typename TI>
__device__ __forceinline__ void load_pair(const TI* p, float& a, float& b);
template <>
__device__ __forceinline__ void load_pair<float>(const float* p, float& a, float& b) {
  const float2 v = *reinterpret_cast<const float2*>(p);
  a = v.x; b = v.y;
}
template <>
__device__ __forceinline__ void load_pair<uint8_t>(const uint8_t* p, float& a, float& b) {
  // u8 -> fp32 without I2F (conversions issue on the quarter-rate XU pipe, 108 of them per operand row):
  // 0x4B000000 | u is the float 2^23 + u, one FADD removes the 2^23
  const uint32_t v = *reinterpret_cast<const uint16_t*>(p);
  a = __uint_as_float(0x4B000000u | (v & 0xffu)) - 8388608.0f;
  b = __uint_as_float(0x4B000000u | (v >> 8)) - 8388608.0f;
}

template <typename TI, bool FP16, bool SILU>
__global__ void __launch_bounds__(kStemThreads, 1)
stem_tc_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_img,
               const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  // pointer arithmetic on the __shared__ array (not an integer round trip) keeps the address space known to the
  // compiler: LDS/STS instead of generic loads for every bias / staging access
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  StemShared* sh = reinterpret_cast<StemShared*>(smem);
  float* sbias = reinterpret_cast<float*>(smem + 1024);
  uint8_t* patch = smem + 1024 + p.bias_bytes;                                         // [kPatchStages] raw patches
  uint8_t* wsm = patch + kPatchStages * kPatchStageBytes;                              // [2][BN x 64] weights
  wsm += (1024u - (smem_u32(wsm) & 1023u)) & 1023u;
  uint8_t* astage = wsm + p.b_bytes;                                                   // [stages][2][128 x 64]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr bool fp16 = FP16;

  {
    const float bscale = epi_half_bias(p.epi) ? 0.5f : 1.0f;
    for (int i = threadIdx.x; i < p.epi.out_c; i += blockDim.x) sbias[i] = p.epi.bias[i] * bscale;
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_w);
    for (int i = 0; i < kStemStages; ++i) { mbar_init(&sh->full[i], kStemBuilders); mbar_init(&sh->empty[i], 1); }
    for (int i = 0; i < kStemAcc; ++i) { mbar_init(&sh->tmem_full[i], 1); mbar_init(&sh->tmem_empty[i], 128); }
    mbar_init(&sh->w_full, 1);
    for (int i = 0; i < kPatchStages; ++i) { mbar_init(&sh->patch_full[i], 1); mbar_init(&sh->patch_empty[i], kStemBuilders); }
    tma_prefetch_desc(&map_img);
    fence_barrier_init();
  }
  if (warp == kWarpMma) {
    tmem_alloc(&sh->tmem_base, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch, bias copy) touches only
  // constants and overlaps the tail of the previous kernel; activations are read/written after the wait
  pdl_launch_dependents();
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp < kWarpMma) {
    // ===================== A builders =====================
    const int m = threadIdx.x & (kStemBuilders - 1);   // operand row = output pixel (hl, wl) of the tile
    const int hl = m >> 4, wl = m & 15;
    const int bgrp = threadIdx.x / kStemBuilders;
    for (int it = bgrp;; it += kStemBuilderGroups) {
      const long long tl = (long long)blockIdx.x + (long long)it * gridDim.x;
      if (tl >= p.num_tiles) break;
      const int stage = it % kStemStages, ps = it % kPatchStages;
      const uint32_t phase = (uint32_t)((it / kStemStages) & 1), pphase = (uint32_t)((it / kPatchStages) & 1);
      // ---- 1. the tile's raw patch [3][20][pitch] (fp32: pitch 36, u8: pitch 48) arrives by TMA
      mbar_wait(&sh->patch_full[ps], pphase);
      const TI* pt = reinterpret_cast<const TI*>(patch + (size_t)ps * kPatchStageBytes);
      // the box starts 16-byte aligned in global memory, kLead pixels left of the patch (TMA faults on
      // a row start that is not a multiple of 16 bytes)
      constexpr int kPitch = sizeof(TI) == 4 ? 40 : 64;
      constexpr int kLead = sizeof(TI) == 4 ? 2 : 14;
      // ---- 2. build operand row m: k = dy*18 + c*6 + dx  <->  patch[c][2*hl+dy][2*wl+dx]
      mbar_wait(&sh->empty[stage], phase ^ 1);
      uint8_t* a0 = astage + (size_t)stage * 32768 + (size_t)m * 128;
      uint32_t wds[4];
      int nw = 0, chunk = 0;
      // two batches of 27 independent shared loads (3 dy x 3 c x 3 pairs) so that the load latency is paid
      // twice per row instead of once per pair
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float fa[27], fb[27];
#pragma unroll
        for (int i = 0; i < 27; ++i) {
          const int dy = half * 3 + i / 9, c = (i % 9) / 3, q = i % 3;
          const TI* src = pt + (c * kPatchRows + 2 * hl + dy) * kPitch + kLead + 2 * wl;
          load_pair<TI>(src + 2 * q, fa[i], fb[i]);
        }
#pragma unroll
        for (int i = 0; i < 27; ++i) {
          // fp16: pixels are fed as x/256 (exact) and the host packs 256*W: BN-folded stem weights (~W/300 for raw
          // 0..255 inputs) would otherwise fall into fp16's subnormal range and lose most of their mantissa
          if (FP16) { fa[i] *= 0.00390625f; fb[i] *= 0.00390625f; }
          wds[nw++] = pack16_t<FP16>(fa[i], fb[i]);
          if (nw == 4) {
            uint8_t* dst = a0 + (size_t)(chunk >> 3) * 16384 + (((chunk & 7) ^ (m & 7)) << 4);
            *reinterpret_cast<uint4*>(dst) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
            nw = 0; ++chunk;
          }
        }
      }
      // 54 words written so far = 13 full chunks + 2 words; pad k = 108..127 with zeros
      {
        uint8_t* dst = a0 + (size_t)(chunk >> 3) * 16384 + (((chunk & 7) ^ (m & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = make_uint4(wds[0], wds[1], 0u, 0u);
        ++chunk;
#pragma unroll
        for (; chunk < 16; ++chunk) {
          uint8_t* d2 = a0 + (size_t)(chunk >> 3) * 16384 + (((chunk & 7) ^ (m & 7)) << 4);
          *reinterpret_cast<uint4*>(d2) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      mbar_arrive(&sh->patch_empty[ps]);      // this thread is done with the raw patch
      fence_proxy_async();                    // generic-proxy stores -> visible to the tensor core
      mbar_arrive(&sh->full[stage]);
    }
  } else if (warp == kWarpMma) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // resident weights: two [BN x 64] K-major chunks
      mbar_arrive_expect_tx(&sh->w_full, p.b_bytes);
      tma_load_2d(&map_w, &sh->w_full, wsm, 0, 0);
      tma_load_2d(&map_w, &sh->w_full, wsm + p.b_bytes / 2, 64, 0);
      mbar_wait(&sh->w_full, 0);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      const uint64_t hi = (uint64_t)p.desc_hi << 32;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        mbar_wait(&sh->tmem_empty[as], aphase ^ 1);
        mbar_wait(&sh->full[stage], phase);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.BNpad);
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) {
          const uint32_t sa = smem_u32(astage + (size_t)stage * 32768 + (size_t)kc * 16384);
          const uint32_t sb = smem_u32(wsm + (size_t)kc * (p.b_bytes / 2));
          uint64_t adesc = hi | (uint64_t)(((sa >> 4) & 0x3FFF) | (1u << 16));
          uint64_t bdesc = hi | (uint64_t)(((sb >> 4) & 0x3FFF) | (1u << 16));
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            umma_f16(d_tmem, adesc, bdesc, p.idesc, (uint32_t)((kc | j) != 0));
            adesc += 2; bdesc += 2;
          }
        }
        umma_commit(&sh->empty[stage]);
        umma_commit(&sh->tmem_full[as]);
        if (++stage == kStemStages) { stage = 0; phase ^= 1; }
        if (++as == kStemAcc) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp == kWarpTma) {
    // ===================== patch producer (TMA) =====================
    if (lane == 0) {
      pdl_wait();
      int ps = 0;
      uint32_t pphase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int b = t / tiles_per_img;
        const int r = t - b * tiles_per_img;
        const int ty = r / p.tiles_w, tx = r - ty * p.tiles_w;
        mbar_wait(&sh->patch_empty[ps], pphase ^ 1);
        if (p.debug & 1) {
          mbar_arrive(&sh->patch_full[ps]);
        } else {
          mbar_arrive_expect_tx(&sh->patch_full[ps], p.patch_tx);
          tma_load_4d(&map_img, &sh->patch_full[ps], patch + (size_t)ps * kPatchStageBytes,
                      2 * tx * kTileW - (sizeof(TI) == 4 ? 4 : 16), 2 * ty * kTileH - 2, 0, b);
        }
        if (++ps == kPatchStages) { ps = 0; pphase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue groups =====================
    pdl_wait();
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int grp = (warp - kWarpEpi) >> 2;
    const int hl = row >> 4, wl = row & 15;
    for (int it = grp;; it += kStemEpiGroups) {
      const long long tl = (long long)blockIdx.x + (long long)it * gridDim.x;
      if (tl >= p.num_tiles) break;
      const int t = (int)tl;
      const int as = it % kStemAcc;
      const uint32_t aphase = (uint32_t)((it / kStemAcc) & 1);
      const int b = t / tiles_per_img;
      const int r = t - b * tiles_per_img;
      const int ty = r / p.tiles_w, tx = r - ty * p.tiles_w;
      const int ho = ty * kTileH + hl, wo = tx * kTileW + wl;
      const bool valid = ho < p.out_h && wo < p.out_w;
      mbar_wait(&sh->tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * p.BNpad);
      const long long pix = ((long long)b * p.out_h + ho) * p.out_w + wo;
      uint16_t* orow = (uint16_t*)p.epi.out + pix * p.epi.out_ld;
      for (int c = 0; c < p.BN; c += 32) {
        const bool two = (c + 16 < p.BN);
        uint32_t ra[16], rb[16];
        tmem_ld_x16(taddr + (uint32_t)c, ra);
        if (two) tmem_ld_x16(taddr + (uint32_t)(c + 16), rb);
        tmem_ld_wait();
        if (valid) {
          epi_tc_chunk<FP16, SILU && !FP16>(p.epi, ra, sbias + c, nullptr, orow + c, b, ho, wo, c);
          if (two) epi_tc_chunk<FP16, SILU && !FP16>(p.epi, rb, sbias + c + 16, nullptr, orow + c + 16, b, ho, wo, c + 16);
        }
      }
      tc_fence_before();
      mbar_arrive(&sh->tmem_empty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();

struct StemLaunch {
  CUtensorMap map_w, map_img;
  StemParams p;
  int grid;
  size_t smem;
};

StemLaunch* stem_alloc() {
  void* p = nullptr;
  if (posix_memalign(&p, 64, sizeof(StemLaunch)) != 0) return nullptr;
  memset(p, 0, sizeof(StemLaunch));
  return reinterpret_cast<StemLaunch*>(p);
}
void stem_free(StemLaunch* p) { free(p); }

int stem_prepare(const void* img, int img_dtype, const void* w, const float* bias, void* out, long long out_ld,
                 int batch, int h, int wd, int out_c, int act, int dtype, StemLaunch* L) {
  YX_REQUIRE(img && w && bias && out, YX_ERR_INVALID_ARG, "stem: null pointer");
  YX_REQUIRE(img_dtype == YX_FP32 || img_dtype == YX_U8, YX_ERR_INVALID_ARG, "stem: image dtype must be fp32 or uint8");
  YX_REQUIRE(dtype == YX_BF16 || dtype == YX_FP16, YX_ERR_INVALID_ARG, "stem: output dtype must be bf16/fp16");
  YX_REQUIRE(batch > 0 && h > 0 && wd > 0 && h % 2 == 0 && wd % 2 == 0, YX_ERR_INVALID_ARG, "stem: H, W must be even");
  YX_REQUIRE(out_c % 16 == 0 && out_c >= 16 && out_c <= 128, YX_ERR_UNSUPPORTED, "stem: out_c=%d (16..128, multiple of 16)", out_c);
  YX_REQUIRE(out_ld % 16 == 0 && out_ld >= out_c && ((uintptr_t)out & 31) == 0, YX_ERR_INVALID_ARG, "stem: out misaligned");
  YX_REQUIRE(((uintptr_t)img & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)bias & 15) == 0, YX_ERR_INVALID_ARG, "stem: pointer alignment");
  EncodeTiledFn encode = get_encode_fn();
  YX_REQUIRE(encode != nullptr, YX_ERR_NO_DEVICE, "cuTensorMapEncodeTiled unavailable (no CUDA driver)");
  StemParams& p = L->p;
  memset(&p, 0, sizeof(p));
  p.img = img; p.img_dtype = img_dtype; p.batch = batch; p.h = h; p.w = wd;
  p.out_h = h / 2; p.out_w = wd / 2;
  p.tiles_w = (p.out_w + kTileW - 1) / kTileW; p.tiles_h = (p.out_h + kTileH - 1) / kTileH;
  p.num_tiles = batch * p.tiles_w * p.tiles_h;
  p.BN = out_c;
  p.BNpad = 32; while (p.BNpad < p.BN) p.BNpad <<= 1;
  p.tmem_cols = (unsigned)(kStemAcc * p.BNpad);
  const unsigned fmt = dtype == YX_BF16 ? 1u : 0u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
  p.desc_hi = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);   // SBO = 8 rows * 128 B, version 1, SWIZZLE_128B
  p.bias_bytes = ((unsigned)out_c * 4u + 1023u) & ~1023u;
  p.b_bytes = 2u * (unsigned)out_c * 128u;                           // two [out_c x 64] chunks, multiples of 2 KB
  EpiParams& e = p.epi;
  e.out_h = p.out_h; e.out_w = p.out_w; e.out_c = out_c; e.act = act; e.dtype = dtype; e.epilogue = YX_EPI_STORE;
  e.bias = bias; e.out = out; e.out_ld = out_ld;
  L->smem = 2048 + p.bias_bytes + (size_t)kPatchStages * kPatchStageBytes + 1024 + p.b_bytes + (size_t)kStemStages * 32768;
  {
    // the NCHW image as a 4-D tensor [W, H, 3, B]; one box = the 40 (u8: 64) x 20 x 3 window holding a tile's patch
    const int es = img_dtype == YX_FP32 ? 4 : 1;
    const cuuint32_t bw = img_dtype == YX_FP32 ? 40 : 64;   // patch is 36 wide; the box starts 16-byte aligned
    YX_REQUIRE(((long long)wd * es) % 16 == 0, YX_ERR_UNSUPPORTED, "stem: image row pitch must be a multiple of 16 bytes");
    cuuint64_t idims[4] = {(cuuint64_t)wd, (cuuint64_t)h, 3, (cuuint64_t)batch};
    cuuint64_t istr[3] = {(cuuint64_t)wd * es, (cuuint64_t)wd * h * es, (cuuint64_t)wd * h * 3 * es};
    cuuint32_t ibox[4] = {bw, (cuuint32_t)kPatchRows, 3, 1};
    cuuint32_t iestr[4] = {1, 1, 1, 1};
    CUresult ri = encode(&L->map_img, img_dtype == YX_FP32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 4,
                         const_cast<void*>(img), idims, istr, ibox, iestr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    YX_REQUIRE(ri == CUDA_SUCCESS, YX_ERR_CUDA, "cuTensorMapEncodeTiled(stem image) failed: %d", (int)ri);
    p.patch_tx = bw * kPatchRows * 3 * es;
    if (const char* e = getenv("YX_STEM_DEBUG")) p.debug = atoi(e);
  }
  const CUtensorMapDataType tdt = dtype == YX_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  cuuint64_t dims[2] = {128, (cuuint64_t)out_c};
  cuuint64_t strides[1] = {256};
  cuuint32_t box[2] = {64, (cuuint32_t)out_c};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&L->map_w, tdt, 2, const_cast<void*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  YX_REQUIRE(r == CUDA_SUCCESS, YX_ERR_CUDA, "cuTensorMapEncodeTiled(stem W) failed: %d", (int)r);
  const int sms = num_sms();
  L->grid = p.num_tiles < sms ? p.num_tiles : sms;
  return YX_OK;
}

int stem_launch(const StemLaunch* L, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
#define YX_STEM_ATTR(T, F) \
  YX_CUDA(cudaFuncSetAttribute(stem_tc_kernel<T, F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
  YX_CUDA(cudaFuncSetAttribute(stem_tc_kernel<T, F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024))
    YX_STEM_ATTR(float, false); YX_STEM_ATTR(float, true); YX_STEM_ATTR(uint8_t, false); YX_STEM_ATTR(uint8_t, true);
#undef YX_STEM_ATTR
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)L->grid);
  cfg.blockDim = dim3((unsigned)kStemThreads);
  cfg.dynamicSmemBytes = L->smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  const bool h16 = L->p.epi.dtype == YX_FP16;
  const bool silu = L->p.epi.act == YX_ACT_SILU;
#define YX_STEM_GO(T, F) do { if (silu) YX_CUDA(cudaLaunchKernelEx(&cfg, stem_tc_kernel<T, F, true>, L->map_w, L->map_img, L->p)); \
    else YX_CUDA(cudaLaunchKernelEx(&cfg, stem_tc_kernel<T, F, false>, L->map_w, L->map_img, L->p)); } while (0)
  if (L->p.img_dtype == YX_FP32) {
    if (h16) YX_STEM_GO(float, true); else YX_STEM_GO(float, false);
  } else {
    if (h16) YX_STEM_GO(uint8_t, true); else YX_STEM_GO(uint8_t, false);
  }
#undef YX_STEM_GO
  return YX_OK;
}

}  // namespace yx
