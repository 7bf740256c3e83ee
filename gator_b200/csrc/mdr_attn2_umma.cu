// MDR 431 x 431 self-attention core (lib/models/vanilla_transformer_encoder.py:36-46), round-2 kernel:
//   out = softmax(q k^T / sqrt(32)) v      per (sample, head), fp16 operands / fp32 accumulate on tcgen05.
//
// Operands arrive PRE-PACKED: per (sample, head) one contiguous 82 944-byte record [Q | K | V], each a 432 x 32 fp16
// image in the tcgen05 no-swizzle core-matrix layout  offset(row, d) = (row/8)*512 + (d/8)*128 + (row%8)*16 + (d%8)*2
// (row 431 = zero padding), written by the layer-chain kernel's q|k|v epilogue (csrc/mdr_chain2_umma.cu) or by
// `qkv_image_kernel` below.  Q and K are K-major operands; the same image of V is the MN-major B operand of P V, so no
// transpose of V exists anywhere.
//
// Persistent, warp-specialised, one CTA per SM, 16 warps (512 threads x 128 registers = the whole register file; a 17th
// warp would be allocated as four - the hardware hands out warps in groups of four):
//   warps 0..13  softmax warps in four warpgroups; warpgroup t owns query tile t (rows 128t..128t+127), thread = query
//                row, and tensor-memory columns [128t, 128t+128): S block (96 keys, fp32) in [0,96), O accumulator in
//                [96,128).  Tile 3 holds rows 384..430 only, so its warps 14 and 15 have no row at all; they are
//   warp 14      the MMA issuer (one lane), and
//   warp 15      the TMA producer: one bulk copy per image into a 2-stage ring (the next (sample, head) loads while this one runs).
// Per key block j (9 blocks of 48 keys) and tile t:   S = Q_t K_j^T  ->  warpgroup t: block max, P = exp2(S c - m) as fp16
// written IN PLACE over S with tcgen05.st (P is the tensor-memory A operand of the next MMA - it never touches shared
// memory)  ->  O_t += P V_j.   S is double-buffered per tile (S0 48 | S1 48 | O 32 columns): Q_t K_{j+1}^T is computed
// while the warpgroup works on block j, and P V_j is issued asynchronously behind it, so a softmax warp only ever waits for
// an MMA that was issued a whole block earlier; there is no block-wide barrier in the loop.
// Softmax is the online form with a lazily updated running maximum: the row maximum only moves (and O is rescaled, in
// tensor memory) when a block exceeds it by more than 2^8 - fp16 holds P <= 2^8 exactly as well as P <= 1.
// fp16 (11-bit significand) instead of the round-1 3-term bf16 split: 2.2e-5 m end-to-end vertex error (CPU emulation,
// tests/probes/attn_precision_cpu.py; plain bf16 would be 1.6e-4 m, over the 1e-4 m budget) at a third of the MMAs and
// none of the residual arithmetic.
#include <cuda_fp16.h>

#include "common.cuh"
#include "umma.cuh"

namespace gator {
namespace {

using namespace umma;

constexpr int V = GATOR_V_COARSE;   // 431
constexpr int VP = 432;
constexpr int DK = 32;
constexpr int E = 64;
constexpr int IMG = VP * DK * 2;          // 27 648 B per image
constexpr int ITEM = 3 * IMG;             // 82 944 B per (sample, head)
constexpr int STAGES = 2;
constexpr int KB = 48;                    // keys per block
constexpr int NBLK = 9;                   // 9 x 48 = 432
constexpr int NT = 512;                   // 16 warps
constexpr int W_MMA = 14, W_TMA = 15;     // the two row-less warps of query tile 3
constexpr int OST_WARP = 32 * DK * 4;      // 4 KB per softmax warp: O rows staged for coalesced stores
constexpr int SMEM = STAGES * ITEM + 14 * OST_WARP;   // 165 888 + 57 344 B dynamic
constexpr float kScaleLog2 = 0.17677669529663687f * 1.4426950408889634f;   // log2(e) / sqrt(32)
constexpr int kEmu = 1;                   // exponentials per group of four evaluated by ex2_poly (see there)
constexpr float kLazy = 8.0f;             // the running maximum moves only when exceeded by 2^8

struct Bars {
  uint64_t kv_full[STAGES], kv_empty[STAGES];
  uint64_t s_full[4][2], p_full[4][2], pv_done[4], o_full[4];
};

// One key block (48 keys) of one query row, single pass over the scores (48 registers).  LAST: the block's last column is
// the zero pad row of K.  s_addr: this thread's lane + first column of the S buffer; o_addr: lane + first column of O.
// 2^x on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f, f in [-0.5, 0.5], cubic minimax polynomial for
// 2^f (relative error 7.5e-5, a sixth of the fp16 rounding P gets anyway), n added into the exponent field.  The MUFU
// unit delivers 16 results per clock and SM and the 2 x 431 x 432 exponentials per sample are this kernel's floor
// (0.33 ms per 4096-sample launch at 100 % MUFU utilisation), so EMU of every four exponentials take this path.
// Measured per 4096-sample launch: EMU = 0 524 us, EMU = 1 512 us, EMU = 2 583 us (the issue slots, not the MUFU unit,
// bind beyond one in four); same error against fp64 in all three (tools/attn2_probe.py).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -100.0f);
  const float fi = x + 12582912.0f;                 // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (fi - 12582912.0f);
  float p = fmaf(0.055171459913253784f, f, 0.2426108568906784f);
  p = fmaf(p, f, 0.6932609677314758f);
  p = fmaf(p, f, 0.9999281167984009f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(fi) << 23));
}

template <bool LAST, int EMU>
__device__ __forceinline__ void softmax_block(uint32_t s_addr, uint32_t o_addr, bool first, bool valid, float& m, float& l,
                                              uint64_t* pv_done, uint32_t pv_parity) {
  uint32_t a[32], b[16];
  tmem_ld32_async(s_addr, a);
  tmem_ld16_async(s_addr + 32, b);
  tmem_ld_wait_dep32(a);
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    m0 = fmax3(m0, __uint_as_float(a[i]), __uint_as_float(a[i + 1]));
    m1 = fmax3(m1, __uint_as_float(a[i + 2]), __uint_as_float(a[i + 3]));
  }
  tmem_ld_wait_dep16(b);
#pragma unroll
  for (int i = 0; i < 12; i += 4) {
    m0 = fmax3(m0, __uint_as_float(b[i]), __uint_as_float(b[i + 1]));
    m1 = fmax3(m1, __uint_as_float(b[i + 2]), __uint_as_float(b[i + 3]));
  }
  m0 = fmax3(m0, __uint_as_float(b[12]), __uint_as_float(b[13]));
  m1 = LAST ? fmaxf(m1, __uint_as_float(b[14])) : fmax3(m1, __uint_as_float(b[14]), __uint_as_float(b[15]));
  const float mb = fmaxf(m0, m1) * kScaleLog2;
  // ---- running maximum (lazy) and, rarely, the rescale of O in tensor memory ----
  if (first) {
    m = mb;
  } else {
    const bool need = valid && (mb > m + kLazy);
    if (__ballot_sync(0xffffffffu, need)) {   // warp-uniform: tcgen05.ld / st are warp-collective
      const float f = need ? ex2_approx(m - mb) : 1.0f;
      if (need) { l *= f; m = mb; }
      mbar_wait(pv_done, pv_parity);          // the previous block's P V (issued asynchronously) has finished accumulating into O
      tc_fence_after();
      uint32_t o[32];
      tmem_ld32_async(o_addr, o);
      tmem_ld_wait_dep32(o);
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
      tmem_st32(o_addr, o);
    }
  }
  // ---- P = exp2(s c - m) -> fp16, written over the first 24 columns of the S buffer; row sum in fp32 ----
  const float nm = -m;
  float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
  uint32_t pk[24];
#pragma unroll
  for (int i = 0; i < 48; i += 4) {
    const uint32_t* s = i < 32 ? a + i : b + (i - 32);
    const float p0 = ex2_approx(fmaf(__uint_as_float(s[0]), kScaleLog2, nm));
    const float x1 = fmaf(__uint_as_float(s[1]), kScaleLog2, nm), x2 = fmaf(__uint_as_float(s[2]), kScaleLog2, nm);
    const float p1 = EMU >= 2 ? ex2_poly(x1) : ex2_approx(x1);
    const float p2 = EMU >= 1 ? ex2_poly(x2) : ex2_approx(x2);
    const float p3 = (LAST && i == 44) ? 0.f : ex2_approx(fmaf(__uint_as_float(s[3]), kScaleLog2, nm));
    l0 += p0; l1 += p1; l2 += p2; l3 += p3;
    pk[i / 2] = pack_f16(p0, p1);
    pk[i / 2 + 1] = pack_f16(p2, p3);
  }
  tmem_st16(s_addr, pk);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(s_addr + 16), "r"(pk[16]), "r"(pk[17]),
               "r"(pk[18]), "r"(pk[19]), "r"(pk[20]), "r"(pk[21]), "r"(pk[22]), "r"(pk[23])
               : "memory");
  l += (l0 + l1) + (l2 + l3);
  tmem_st_wait();
}

template <int EMU>
__global__ void __launch_bounds__(NT, 1) mdr_self_attn2_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, int items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ Bars bars;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;

  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 32) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bars.kv_full[s], 1); mbar_init(&bars.kv_empty[s], 1); }
    for (int t = 0; t < 4; ++t) {
      // warpgroup 3 holds rows 384..511 of which 384..430 exist: warps 2, 3 of it have no row at all and take no part
      mbar_init(&bars.s_full[t][0], 1);
      mbar_init(&bars.s_full[t][1], 1);
      // (a warp may run one block ahead of its warpgroup - S of the next block is ready early - so the "P written"
      //  barrier is per S buffer: arrivals of block g and block g + 1 never mix)
      mbar_init(&bars.p_full[t][0], t < 3 ? 4 : 2);
      mbar_init(&bars.p_full[t][1], t < 3 ? 4 : 2);
      mbar_init(&bars.pv_done[t], 1);
      mbar_init(&bars.o_full[t], 1);
    }
    mbar_init_fence();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int n_my = (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // items of this CTA: blockIdx.x + i * gridDim.x

  if (warp == W_TMA) {
    // ===== TMA producer =====
    if (elect_one()) {
      for (int i = 0; i < n_my; ++i) {
        const int s = i % STAGES;
        if (i >= STAGES) mbar_wait(&bars.kv_empty[s], ((i / STAGES) - 1) & 1);
        const uint8_t* src = img + (size_t)(blockIdx.x + (size_t)i * gridDim.x) * ITEM;
        uint8_t* dst = smem + s * ITEM;
        mbar_arrive_expect_tx(&bars.kv_full[s], ITEM);
#pragma unroll
        for (int c = 0; c < 3; ++c) bulk_copy_g2s(dst + c * IMG, src + c * IMG, IMG, &bars.kv_full[s]);
      }
    }
  } else if (warp == W_MMA) {
    // ===== MMA issuer =====
    constexpr uint32_t id_s = idesc_f16(128, KB), id_pv = idesc_f16(128, DK, 1);
    // Blocks are numbered globally, g = item * NBLK + j; block g uses S buffer g & 1 of its tile (NBLK is odd, so the buffer of a
    // (sample, head)'s first block alternates from one item to the next).
    auto issue_qk = [&](uint32_t stage_base, int t, int j, int buf) {   // S_t[buf] = Q_t K_j^T (2 k-steps of 16)
      const uint32_t q0 = stage_base + t * (16 * 512), k0 = stage_base + IMG + j * (KB / 8) * 512;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
        mma_bf16(tmem + t * 128 + buf * KB, smem_desc(q0 + ks * 256, 128, 512), smem_desc(k0 + ks * 256, 128, 512), id_s, ks);
    };
    auto issue_pv = [&](uint32_t stage_base, int t, int j, int buf) {   // O_t (+)= P V_j: A = P in tensor memory (first 24 columns of S_t[buf]), B = V MN-major
      const uint32_t v0 = stage_base + 2 * IMG + j * (KB / 8) * 512;
#pragma unroll
      for (int ks = 0; ks < KB / 16; ++ks)
        mma_ts(tmem + t * 128 + 2 * KB, tmem + t * 128 + buf * KB + ks * 8, smem_desc(v0 + ks * 1024, 512, 128), id_pv, (j | ks) != 0);
    };
    uint32_t ph_p[2] = {0, 0};   // parity of the next p_full completion per S buffer (the four tiles are served in turn; serving
                                 // them out of order by polling the barriers was measured 7 % slower: the polling warp takes issue slots)
    for (int i = 0; i < n_my; ++i) {
      const int s = i % STAGES;
      const uint32_t sb = smem_u32(smem + s * ITEM);
      if (i == 0) {
        mbar_wait(&bars.kv_full[s], 0);
        if (elect_one()) {
          for (int t = 0; t < 4; ++t) {
            issue_qk(sb, t, 0, 0); mma_commit(&bars.s_full[t][0]);
            issue_qk(sb, t, 1, 1); mma_commit(&bars.s_full[t][1]);
          }
        }
        __syncwarp();
      }
      for (int j = 0; j < NBLK; ++j) {
        const int buf = (i * NBLK + j) & 1;
        uint32_t sb_next = 0;
        if (j >= NBLK - 2 && i + 1 < n_my) {          // the last two blocks are followed by the first two of the next (sample, head)
          const int s1 = (i + 1) % STAGES;
          mbar_wait(&bars.kv_full[s1], ((i + 1) / STAGES) & 1);
          sb_next = smem_u32(smem + s1 * ITEM);
        }
        for (int t = 0; t < 4; ++t) {
          mbar_wait(&bars.p_full[t][buf], ph_p[buf]);
          tc_fence_after();
          if (elect_one()) {
            issue_pv(sb, t, j, buf);
            mma_commit(&bars.pv_done[t]);
            if (j == NBLK - 1) {
              mma_commit(&bars.o_full[t]);
              if (t == 3) mma_commit(&bars.kv_empty[s]);   // every MMA that reads this stage has been issued
            }
            // the S buffer just consumed takes the block after next (in-order tensor pipe: P V has read P before it is overwritten)
            if (j + 2 < NBLK) {
              issue_qk(sb, t, j + 2, buf);
              mma_commit(&bars.s_full[t][buf]);
            } else if (sb_next) {
              issue_qk(sb_next, t, j + 2 - NBLK, buf);
              mma_commit(&bars.s_full[t][buf]);
            }
          }
          __syncwarp();
        }
        ph_p[buf] ^= 1;
      }
    }
  } else {
    // ===== softmax warpgroups =====
    const int t = warp >> 2;                        // query tile
    const int row = (warp & 3) * 32 + lane;         // row in tile = tensor-memory lane
    const int q = t * 128 + row;
    const bool valid = q < V;
    {
      const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
      const uint32_t s_base = tmem + lane_addr + t * 128, o_addr = s_base + 2 * KB;
      uint32_t ph_s = 0, ph_o = 0, n_pv = 0;          // ph_s bit buf: parity of the next s_full[t][buf] completion; n_pv: P V groups so far
      for (int i = 0; i < n_my; ++i) {
        const int item = blockIdx.x + i * gridDim.x;
        float m = 0.f, l = 0.f;
#pragma unroll 1
        for (int j = 0; j < NBLK; ++j) {
          // buffer of block j of item i: blocks alternate buffers ACROSS items too (NBLK is odd)
          const int buf = (i * NBLK + j) & 1;
          mbar_wait(&bars.s_full[t][buf], (ph_s >> buf) & 1);
          ph_s ^= 1u << buf;
          tc_fence_after();
          const uint32_t s_addr = s_base + buf * KB;
          if (j < NBLK - 1) softmax_block<false, EMU>(s_addr, o_addr, j == 0, valid, m, l, &bars.pv_done[t], (n_pv - 1) & 1);
          else softmax_block<true, EMU>(s_addr, o_addr, false, valid, m, l, &bars.pv_done[t], (n_pv - 1) & 1);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars.p_full[t][buf]);
          ++n_pv;
        }
        // ---- O out: thread = row, 32 fp32 = one 128-byte line of out[b, q, h*32 .. h*32+31] ----
        mbar_wait(&bars.o_full[t], ph_o);
        ph_o ^= 1;
        tc_fence_after();
        uint32_t o[32];
        tmem_ld32_async(o_addr, o);
        tmem_ld_wait_dep32(o);
        {
          // row-per-lane -> (4 rows x 128 contiguous bytes) per store instruction, through a warp-private XOR-swizzled
          // shared-memory tile (a lane-per-row store touches 32 half-used sectors per instruction)
          const float inv = 1.0f / l;
          uint8_t* st = smem + STAGES * ITEM + warp * OST_WARP;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4*>(st + lane * 128 + ((c ^ (lane & 7)) * 16)) =
                make_float4(__uint_as_float(o[4 * c]) * inv, __uint_as_float(o[4 * c + 1]) * inv, __uint_as_float(o[4 * c + 2]) * inv,
                            __uint_as_float(o[4 * c + 3]) * inv);
          __syncwarp();
          const int c = lane & 7;
          const int q0 = t * 128 + (warp & 3) * 32;         // first query row of this warp
          float* obase = out + ((size_t)(item >> 1) * V) * E + (item & 1) * DK + c * 4;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int r = k * 4 + (lane >> 3);
            const float4 val = *reinterpret_cast<const float4*>(st + r * 128 + ((c ^ (r & 7)) * 16));
            if (q0 + r < V) *reinterpret_cast<float4*>(obase + (size_t)(q0 + r) * E) = val;
          }
          __syncwarp();
        }
        // the next item's first P V (accumulate = 0) overwrites O only after this thread's p_full arrival for its block 0
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// fp32 qkv rows (nb*431, 192) -> per (sample, head) [Q | K | V] fp16 images.  One thread per 16-byte chunk
// (8 consecutive d of one row of one tensor); row 431 is written as zeros.
__global__ void __launch_bounds__(256) qkv_image_kernel(const float* __restrict__ qkv, uint8_t* __restrict__ img, int nb) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)nb * 2 * 3 * VP * 4;
  if (idx >= total) return;
  const int kc = (int)(idx & 3);
  long long r = idx >> 2;
  const int row = (int)(r % VP);
  r /= VP;
  const int which = (int)(r % 3);
  r /= 3;
  const int h = (int)(r & 1);
  const long long b = r >> 1;
  uint4 o = make_uint4(0, 0, 0, 0);
  if (row < V) {
    const float4* src = reinterpret_cast<const float4*>(qkv + ((size_t)b * V + row) * (3 * E) + which * E + h * DK + kc * 8);
    const float4 x = __ldg(src), y = __ldg(src + 1);
    o = make_uint4(pack_f16(x.x, x.y), pack_f16(x.z, x.w), pack_f16(y.x, y.y), pack_f16(y.z, y.w));
  }
  uint8_t* dst = img + ((size_t)(b * 2 + h) * 3 + which) * IMG + (row >> 3) * 512 + kc * 128 + (row & 7) * 16;
  *reinterpret_cast<uint4*>(dst) = o;
}

}  // namespace

size_t self_attn2_image_bytes(int nb) { return (size_t)nb * 2 * ITEM; }

int launch_qkv_image(const float* qkv, void* img, int nb, cudaStream_t stream) {
  const long long total = (long long)nb * 2 * 3 * VP * 4;
  qkv_image_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(qkv, static_cast<uint8_t*>(img), nb);
  return check_launch("qkv_image");
}

// img: self_attn2_image_bytes(nb) bytes of [Q | K | V] records; out (nb*431, 64) fp32
int launch_self_attn2(const void* img, float* out, int nb, cudaStream_t stream) {
  static DeviceOnce attr_once;
  static int num_sms[64];
  int dev = 0;
  cudaGetDevice(&dev);
  GATOR_TRY(attr_once.run("mdr_self_attn2", [&](int d) -> cudaError_t {
    GATOR_CUDA_OK(cudaFuncSetAttribute(mdr_self_attn2_kernel<kEmu>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    return cudaDeviceGetAttribute(&num_sms[d & 63], cudaDevAttrMultiProcessorCount, d);
  }));
  const int items = nb * 2;
  const int sms = num_sms[dev & 63] > 0 ? num_sms[dev & 63] : 148;
  const int grid = items < sms ? items : sms;
  mdr_self_attn2_kernel<kEmu><<<grid, NT, SMEM, stream>>>(static_cast<const uint8_t*>(img), out, items);
  return check_launch("mdr_self_attn2");
}

}  // namespace gator
