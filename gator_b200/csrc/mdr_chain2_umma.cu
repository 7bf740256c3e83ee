// Fused MDR layer "chain", round-2 kernel (lib/models/MDR.py:140-153 minus the 431x431 softmax core) - everything that is
// row-wise between two self-attention cores, for 128 token rows per tile:
//
//   x   = x3_prev + selfatt_prev.linears[3](att_prev) + b          (vanilla_transformer_encoder.py:94)   [layers 1,2]
//   q   = LayerNorm1(x) Wq^T                                        (MDR.py:65, :37)
//   a   = softmax_J(q k^T / sqrt(32)) v     per head, k/v of the sample's J joints                       (:40-43)
//   x  += a Wproj^T + b                                             (:44, :66)
//   x  += fc2(GELU(fc1(LayerNorm2(x))))                             (:68)
//   x3  = a2 (x - mean) / (std_unbiased + 1e-6) + b2                (vanilla_transformer_encoder.py:31-34)
//   q|k|v = x3 [Wq;Wk;Wv]^T + b                                     (:88-90)   -> x3 (fp32) and fp16 operand images
//
// What changed against the round-1 kernel (A images in shared memory, 18 block-wide MMA round trips per tile), which spent
// two thirds of its cycles with no eligible warp:
//  * every A operand lives in TENSOR MEMORY (tcgen05.mma [d], [a], b-desc): the row owner writes its bf16 hi / lo halves
//    with tcgen05.st into its own lane - no shared-memory A images, no fence.proxy.async, no bank conflicts;
//  * the residual stream x lives in tensor memory too and IS the accumulator of the three residual GEMMs
//    (linears[3], proj, fc2 accumulate straight into it; their biases are added when x is next read);
//  * GELU(fc1) is converted IN PLACE (32 fp32 columns -> 16 hi + 16 lo columns) and read back as the A operand of fc2;
//  * 7 MMA round trips per tile instead of 18 (fc1 / fc2 in halves of 128 columns, q|k|v as one group);
//  * the 14 weight units (16 KB each) stream through a 5-slot ring with cp.async.bulk + mbarriers, issued a whole compute
//    phase ahead; 8 warps own the rows (two threads per row) and lane 0 of warp 0 doubles as MMA issuer and TMA producer
//    (256 threads x 128 registers x 2 CTAs = the whole register file; a 9th warp would be allocated as four); the only
//    block-wide synchronisation in the tile loop is the mbarrier pair around each MMA group; persistent CTAs, two per SM
//    (256 tensor-memory columns each);
//  * q|k|v leave as the fp16 operand images the self-attention kernel (csrc/mdr_attn2_umma.cu) loads with bulk copies.
#include <cuda_fp16.h>

#include "common.cuh"
#include "umma.cuh"

namespace gator {
namespace {

using namespace umma;

constexpr int V = GATOR_V_COARSE;   // 431
constexpr int VP = 432;
constexpr int E = 64;
constexpr int DK = 32;
constexpr int MAXJ = 32;
constexpr int UNIT_IMG = 64 * 64 * 2;          // 8 KB: one 64x64 bf16 image
constexpr int UNIT_BYTES = 2 * UNIT_IMG;       // hi | lo
constexpr int SLOTS = 5;
constexpr int NCOMP = 8;                       // warps: all of them own rows
constexpr int NT = NCOMP * 32;
constexpr int QKV_IMG = VP * DK * 2;           // one fp16 operand image of the self-attention kernel
// tensor-memory columns (256 per CTA)
constexpr int C_X = 0;        // residual stream / accumulator of the residual GEMMs (fp32, 64 columns)
constexpr int C_W = 64;       // work region (128 columns): q, fc1 halves (GELU'd in place), head
constexpr int C_A = 192;      // A operand of K = 64: hi columns [192,224), lo columns [224,256)
// parameter arrays (index into Chain2Params::prm), same order as the round-1 kernel
enum { P_SO_B = 0, P_N1W, P_N1B, P_PROJ_B, P_N2W, P_N2B, P_FC1_B, P_FC2_B, P_CLN_A, P_CLN_B, P_QKV_B };
// weight units of the layer blob in storage order (gator_b200/models/MDR.py: pack) ...
enum { U_SO = 0, U_Q = 1, U_PROJ = 2, U_FC1 = 3, U_FC2 = 7, U_QKV = 11 };
// ... and in the order the kernel consumes them
__constant__ int kUnitOrder[14] = {U_SO, U_Q, U_PROJ, U_FC1, U_FC1 + 1, U_FC2, U_FC2 + 1, U_FC1 + 2, U_FC1 + 3, U_FC2 + 2, U_FC2 + 3, U_QKV, U_QKV + 1, U_QKV + 2};

struct Chain2Params {
  const float* x_in;      // (nb*431, 64): layer 0: embedded vertices; layers 1,2: x3 of the previous layer
  const float* att_in;    // (nb*431, 64) self-attention output of the previous layer, or null (layer 0)
  const float* kv;        // (nb*J, 128) this layer's cross-attention K | V
  const uint8_t* blob;    // 14 x UNIT_BYTES packed bf16 weight images (unit 0 = previous layer's linears[3])
  const float* prm[11];
  float* x3_out;          // (nb*431, 64)
  float* qkv_out;         // optional (nb*431, 192) fp32 q|k|v
  uint8_t* img_out;       // optional fp16 [Q | K | V] operand images, one 3 x 27 648-byte record per (sample, head)
  float* hd_out;          // non-null: FINAL pass - hd = (x_in + att_in Wo^T + b_o) W_head^T + b_head (nb*431, 28), folded;
                          // blob = [head Wh (28 rows zero-padded to 64), Wh Wo, 64 floats of folded bias Wh b_o + b_h]
  int J;
  long long rows_total;   // nb * 431
  int ntiles;
  // layer 0 with x_in == null: the vertex embedding (MDR.py:127-134 with the constants folded) is computed on the fly,
  //   x[b, v, :] = VF_CONST[v, :] + W_v[:, 3:6] . pose3d[b, vj[v]] (/ 1000 when pose3d is in millimetres)
  ChainEmbed emb;
};

struct Bars {
  uint64_t w_full[SLOTS];        // weight unit landed (TMA transaction bytes)
  uint64_t a_ready, d_ready;     // A operands written by all 8 warps / MMA group complete (which also frees its ring slots)
};
// MMA groups of a tile, in order
enum Step { S_SO, S_Q, S_PROJ, S_FC1A, S_MID, S_FC2B, S_QKV, S_HEAD };

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// 32 fp32 values -> 16 packed bf16x2 "hi" registers and 16 packed residual registers (k ascending, even k in the low half)
__device__ __forceinline__ void split32(const float* v, uint32_t* hi, uint32_t* lo) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t h = pack_bf16(v[2 * i], v[2 * i + 1]);
    hi[i] = h;
    lo[i] = pack_bf16(v[2 * i] - bf16_lo_f(h), v[2 * i + 1] - bf16_hi_f(h));
  }
}

template <int JT>
__global__ void __launch_bounds__(NT, 2) mdr_chain2_kernel(Chain2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];   // [SLOTS][UNIT_BYTES] weight ring | K|V [2 samples][J][128] fp32
  __shared__ Bars bars;
  __shared__ uint32_t tmem_slot;
  __shared__ float2 xch_buf[2][128][2];               // LayerNorm partial statistics, [parity][row][column half]
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const bool final_pass = p.hd_out != nullptr;
  const bool has_so = p.att_in != nullptr;
  const int J = JT ? JT : p.J;
  float* skv = reinterpret_cast<float*>(smem + SLOTS * UNIT_BYTES);

  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  if (tid == 32) {
    for (int s = 0; s < SLOTS; ++s) mbar_init(&bars.w_full[s], 1);
    mbar_init(&bars.a_ready, NCOMP);
    mbar_init(&bars.d_ready, 1);
    mbar_init_fence();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int n_my = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  // weight units per tile, in consumption order: final pass: so, head; layer pass: [so,] q, proj, fc1 x2, fc2 x2, fc1 x2, fc2 x2, q, k, v
  const int first_unit = final_pass ? 0 : (has_so ? 0 : 1);
  const int units_per_tile = final_pass ? 2 : 14 - first_unit;

  // ---- leader warp 0: TMA producer + MMA issuer (executed by the whole warp, instructions issued by one elected lane, so
  //      that descriptors and addresses stay in uniform registers) ----
  const int total_units = n_my * units_per_tile;
  int u_use = 0, u_load = 0;                         // weight units consumed by issued MMAs / requested from L2
  uint32_t ph_a = 0;
  auto refill = [&]() {                              // every unit below u_use has been read (d_ready): its slot is free
    while (u_load < total_units && u_load - u_use < SLOTS) {
      const int s = u_load % SLOTS, k = u_load % units_per_tile;
      const int unit = final_pass ? k : kUnitOrder[first_unit + k];
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars.w_full[s], UNIT_BYTES);
        bulk_copy_g2s(smem + s * UNIT_BYTES, p.blob + (size_t)unit * UNIT_BYTES, UNIT_BYTES, &bars.w_full[s]);
      }
      ++u_load;
    }
  };
  // one 64x64 unit: D[dcol .. dcol+64) (+)= A . W^T with the 3-term split; A hi k-step ks at a_hi + (ks>>1)*a_grp2 + (ks&1)*8
  // (8 columns each), lo a_lo_off columns further
  auto unit_mma = [&](uint32_t dcol, bool accumulate, uint32_t a_hi, uint32_t a_lo_off, uint32_t a_grp2) {
    constexpr uint32_t idesc = idesc_bf16(128, 64);
    const int s = u_use % SLOTS;
    mbar_wait(&bars.w_full[s], (u_use / SLOTS) & 1);
    const uint32_t w0 = smem_u32(smem + s * UNIT_BYTES);
    if (elect_one()) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t ah = tmem + a_hi + (ks >> 1) * a_grp2 + (ks & 1) * 8, al = ah + a_lo_off;
        const uint64_t wh = smem_desc(w0 + ks * 256, 128, 1024), wl = smem_desc(w0 + UNIT_IMG + ks * 256, 128, 1024);
        mma_ts(tmem + dcol, al, wh, idesc, (accumulate || ks > 0) ? 1u : 0u);
        mma_ts(tmem + dcol, ah, wl, idesc, 1);
        mma_ts(tmem + dcol, ah, wh, idesc, 1);
      }
    }
    __syncwarp();
    ++u_use;
  };
  // A operand region C_A: hi k-step s at C_A + 8 s, lo 32 columns further
  auto from_a = [&](uint32_t dcol, bool acc) { unit_mma(dcol, acc, C_A, 32, 16); };
  // GELU'd fc1 half in C_W: 32-column groups [hi 16 | lo 16]; fc2 quarter qh (0/1 within the half) = groups 2qh, 2qh+1
  auto from_w = [&](int qh) { unit_mma(C_X, true, C_W + qh * 64, 16, 32); };
  auto issue = [&](int step) {                       // leader only
    mbar_wait(&bars.a_ready, ph_a);
    ph_a ^= 1;
    tc_fence_after();
    switch (step) {
      case S_SO: from_a(C_X, true); break;                                                    // x += att Wo^T
      case S_Q: from_a(C_W, false); break;                                                    // q = LN1(x) Wq^T
      case S_PROJ: from_a(C_X, true); break;                                                  // x += a Wproj^T
      case S_FC1A: from_a(C_W, false); from_a(C_W + 64, false); break;                        // fc1 quarters 0, 1
      case S_MID: from_w(0); from_w(1); from_a(C_W, false); from_a(C_W + 64, false); break;   // x += fc2 half 0; fc1 quarters 2, 3
      case S_FC2B: from_w(0); from_w(1); break;                                               // x += fc2 half 1
      case S_QKV: from_a(C_X, false); from_a(C_W, false); from_a(C_W + 64, false); break;     // q | k | v into columns [0, 192)
      default: from_a(C_X, false); unit_mma(C_X, true, C_W, 16, 32); break;                   // S_HEAD: hd = x3 Wh^T + att (Wh Wo)^T
    }
    if (elect_one()) mma_commit(&bars.d_ready);
    __syncwarp();
  };
  const bool leader = warp == 0;
  if (leader) refill();
  {
    // ===== compute warps: two threads per row (one 32-column half each) =====
    const int ch = warp >> 2;                        // column half = head
    const int row = (warp & 3) * 32 + lane;          // row in tile = tensor-memory lane
    const int c0 = ch * 32;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t t_x = tmem + lane_addr + C_X + c0;              // my half of x
    const uint32_t t_ahi = tmem + lane_addr + C_A + ch * 16, t_alo = t_ahi + 32;
    const int pair_id = 2 + (warp & 3);
    uint32_t ph_d = 0;
    int ln_count = 0;
    auto submit = [&](int step) {                    // my part of the A operands is written: hand the MMA group over
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.a_ready);
      if (leader) issue(step);
      __syncwarp();
    };
    auto await = [&]() {
      mbar_wait(&bars.d_ready, ph_d);
      ph_d ^= 1;
      tc_fence_after();
      if (leader) refill();
      __syncwarp();
    };
    auto ld32f = [&](uint32_t taddr, float* v) {
      uint32_t r[32];
      tmem_ld32_async(taddr, r);
      tmem_ld_wait_dep32(r);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
    };
    auto write_a = [&](const float* v) {             // my 32 k-values of a K = 64 A operand
      uint32_t hi[16], lo[16];
      split32(v, hi, lo);
      tmem_st16(t_ahi, hi);
      tmem_st16(t_alo, lo);
    };
    auto prm4 = [&](int which, int off) { return __ldg(reinterpret_cast<const float4*>(p.prm[which] + off)); };
    // LayerNorm statistics of the 64-wide row from its two 32-wide halves (two-pass per half, parallel-variance combine)
    auto stats = [&](const float* xr, float& mean, float& m2) {
      float2* base = &xch_buf[ln_count & 1][row][0];
      ++ln_count;
      float m4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 32; ++i) m4[i & 3] += xr[i];
      const float m = ((m4[0] + m4[1]) + (m4[2] + m4[3])) * (1.0f / 32);
      float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 32; ++i) { const float d = xr[i] - m; q4[i & 3] = fmaf(d, d, q4[i & 3]); }
      const float q = (q4[0] + q4[1]) + (q4[2] + q4[3]);
      base[ch] = make_float2(m, q);
      named_sync(pair_id, 64);
      const float2 o = base[1 - ch];
      const float dm = m - o.x;
      mean = 0.5f * (m + o.x);
      m2 = (q + o.y) + dm * dm * 16.0f;
    };

    for (int it = 0; it < n_my; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const long long flat = (long long)tile * 128 + row;
      const bool valid = flat < p.rows_total;
      const int b_first = (int)(((long long)tile * 128) / V);
      const size_t grow = valid ? (size_t)flat : (size_t)tile * 128;   // rows past the end behave like the tile's first row (never stored)
      const int b = (int)(grow / V);
      const int vert = (int)(grow - (size_t)b * V);
      float x[32], v[32];
      // ---- K|V of the (at most two) samples the tile's rows belong to -> shared memory ----
      if (!final_pass) {
        named_sync(1, NCOMP * 32);                   // everyone is done with the previous tile's K|V
        const long long nb_total = p.rows_total / V;
        const int ns = (b_first + 1 < nb_total) ? 2 : 1;
        const float* src = p.kv + (size_t)b_first * J * 128;
        const uint32_t dst = smem_u32(skv);
        for (int i = tid; i < ns * J * 32; i += NCOMP * 32) cp_async16(dst + i * 16, src + i * 4);
        cp_async_commit();
      }
      // ---- x -> tensor memory; layers 1, 2 / final: A = att, x += att Wo^T on the tensor core ----
      if (p.x_in) {
        const float4* src = reinterpret_cast<const float4*>(p.x_in + grow * E + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = valid ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          x[4 * i] = t.x; x[4 * i + 1] = t.y; x[4 * i + 2] = t.z; x[4 * i + 3] = t.w;
        }
      } else {
        const float* pj = p.emb.pose3d + ((size_t)b * J + __ldg(p.emb.vj + vert)) * 3;
        float px = __ldg(pj), py = __ldg(pj + 1), pz = __ldg(pj + 2);
        if (!p.emb.metres) { px = px / 1000.0f; py = py / 1000.0f; pz = pz / 1000.0f; }
        const float4* vc = reinterpret_cast<const float4*>(p.emb.vconst + (size_t)vert * E + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = __ldg(vc + i);
          const float r4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float* w = p.emb.w3 + (c0 + 4 * i + u) * 3;
            x[4 * i + u] = valid ? r4[u] + fmaf(__ldg(w), px, fmaf(__ldg(w + 1), py, __ldg(w + 2) * pz)) : 0.f;
          }
        }
      }
      if (tid == 0 && it + 1 < n_my) {
        // the next tile's rows of x (and att) are 128 consecutive 256-byte rows: two bulk prefetches bring them from HBM
        // into L2 while this tile computes, so the loads at the top of the next iteration do not pay the DRAM latency
        // (issued before the FINAL pass branches off: that pass is little more than these loads - without the prefetch it
        //  ran at 3.1 TB/s with 77 % of its stall samples on the scoreboard, profiles/r02_ncu_top_kernels.txt)
        const long long r0 = (long long)(tile + gridDim.x) * 128;
        const long long nr = (p.rows_total - r0 < 128) ? p.rows_total - r0 : 128;
        if (nr > 0) {
          if (p.x_in) bulk_prefetch_l2(p.x_in + r0 * E, (uint32_t)(nr * E * 4));
          if (has_so) bulk_prefetch_l2(p.att_in + r0 * E, (uint32_t)(nr * E * 4));
        }
      }
      auto store_x = [&]() {                         // x (registers) -> my half of the accumulator columns
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(x[i]);
        tmem_st32(t_x, r);
      };
      if (final_pass) {
        // FINAL pass: hd = (x3 + att Wo^T + b_o) Wh^T + b_h = x3 Wh^T + att (Wh Wo)^T + (Wh b_o + b_h) - the composite
        // weight and bias are folded at pack time (gator_b200/models/MDR.py), so the pass is ONE MMA group (two K = 64
        // units into the same accumulator) and x is never materialised: [motion_linear | bias_linear | scale_linear] (MDR.py:156-162)
        const float4* src = reinterpret_cast<const float4*>(p.att_in + grow * E + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = valid ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
        write_a(x);                                  // x3 -> A region (K = 64)
        {
          uint32_t hi[16], lo[16];                   // att -> work columns, 32-column group `ch`: [hi 16 | lo 16]
          split32(v, hi, lo);
          const uint32_t taddr = tmem + lane_addr + C_W + ch * 32;
          tmem_st16(taddr, hi);
          tmem_st16(taddr + 16, lo);
        }
        submit(S_HEAD);
        await();
        if (ch == 0) {
          ld32f(tmem + lane_addr + C_X, v);
          if (valid) {
            const float* hb = reinterpret_cast<const float*>(p.blob + 2 * UNIT_BYTES);   // folded bias behind the two units
            float* dst = p.hd_out + grow * 28;
#pragma unroll
            for (int i = 0; i < 28; i += 4) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(hb + i));
              *reinterpret_cast<float4*>(dst + i) = make_float4(v[i] + bb.x, v[i + 1] + bb.y, v[i + 2] + bb.z, v[i + 3] + bb.w);
            }
          }
        }
        continue;
      }
      store_x();
      if (has_so) {
        const float4* src = reinterpret_cast<const float4*>(p.att_in + grow * E + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = valid ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
        write_a(v);
        submit(S_SO);
        await();
        // the bias of a residual GEMM is added when x is next read, and the sum written back before the next accumulation
        ld32f(t_x, x);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = prm4(P_SO_B, c0 + 4 * i);
          x[4 * i] += t.x; x[4 * i + 1] += t.y; x[4 * i + 2] += t.z; x[4 * i + 3] += t.w;
        }
        store_x();
      }
      // ---- LayerNorm1 -> q ----
      {
        float mean, m2;
        stats(x, mean, m2);
        const float rstd = rsqrtf(m2 * (1.0f / E) + 1e-5f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 w = prm4(P_N1W, c0 + 4 * i), bb = prm4(P_N1B, c0 + 4 * i);
          v[4 * i] = (x[4 * i] - mean) * rstd * w.x + bb.x; v[4 * i + 1] = (x[4 * i + 1] - mean) * rstd * w.y + bb.y;
          v[4 * i + 2] = (x[4 * i + 2] - mean) * rstd * w.z + bb.z; v[4 * i + 3] = (x[4 * i + 3] - mean) * rstd * w.w + bb.w;
        }
        write_a(v);
        submit(S_Q);
      }
      cp_async_wait_all();
      named_sync(1, NCOMP * 32);                     // K|V staged (every thread's copies landed)
      await();
      // ---- cross attention against the sample's J joints: this thread owns head `ch` of its row ----
      {
        float q[DK];
        ld32f(tmem + lane_addr + C_W + c0, q);
        constexpr int JU = JT ? JT : MAXJ;
        // a tile of 128 consecutive rows spans at most two samples (431 rows each): both K|V blocks are in shared memory
        const float* kvb = skv + (b - b_first) * J * 128;
        {
          float s[JU];
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < JU; ++j) {
            if (JT || j < J) {
              const float4* kr = reinterpret_cast<const float4*>(kvb + j * 128 + c0);
              float a0_ = 0.f, a1_ = 0.f, a2_ = 0.f, a3_ = 0.f;
#pragma unroll
              for (int d4 = 0; d4 < DK / 4; ++d4) {
                const float4 kk = kr[d4];
                a0_ = fmaf(q[4 * d4], kk.x, a0_); a1_ = fmaf(q[4 * d4 + 1], kk.y, a1_);
                a2_ = fmaf(q[4 * d4 + 2], kk.z, a2_); a3_ = fmaf(q[4 * d4 + 3], kk.w, a3_);
              }
              s[j] = ((a0_ + a1_) + (a2_ + a3_)) * 0.17677669529663687f;
              mx = fmaxf(mx, s[j]);
            }
          }
          float l = 0.f;
#pragma unroll
          for (int j = 0; j < JU; ++j)
            if (JT || j < J) { s[j] = expf(s[j] - mx); l += s[j]; }
          const float inv = 1.0f / l;
#pragma unroll
          for (int d = 0; d < DK; ++d) v[d] = 0.f;
#pragma unroll
          for (int j = 0; j < JU; ++j) {
            if (JT || j < J) {
              const float pj = s[j] * inv;
              const float4* vr = reinterpret_cast<const float4*>(kvb + j * 128 + E + c0);
#pragma unroll
              for (int d4 = 0; d4 < DK / 4; ++d4) {
                const float4 vv = vr[d4];
                v[4 * d4] = fmaf(pj, vv.x, v[4 * d4]); v[4 * d4 + 1] = fmaf(pj, vv.y, v[4 * d4 + 1]);
                v[4 * d4 + 2] = fmaf(pj, vv.z, v[4 * d4 + 2]); v[4 * d4 + 3] = fmaf(pj, vv.w, v[4 * d4 + 3]);
              }
            }
          }
        }
        write_a(v);
        submit(S_PROJ);
      }
      await();
      // ---- LayerNorm2 -> MLP ----
      {
        ld32f(t_x, x);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = prm4(P_PROJ_B, c0 + 4 * i);
          x[4 * i] += t.x; x[4 * i + 1] += t.y; x[4 * i + 2] += t.z; x[4 * i + 3] += t.w;
        }
        store_x();                                   // fc2 accumulates onto x + proj bias
        float mean, m2;
        stats(x, mean, m2);
        const float rstd = rsqrtf(m2 * (1.0f / E) + 1e-5f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 w = prm4(P_N2W, c0 + 4 * i), bb = prm4(P_N2B, c0 + 4 * i);
          v[4 * i] = (x[4 * i] - mean) * rstd * w.x + bb.x; v[4 * i + 1] = (x[4 * i + 1] - mean) * rstd * w.y + bb.y;
          v[4 * i + 2] = (x[4 * i + 2] - mean) * rstd * w.z + bb.z; v[4 * i + 3] = (x[4 * i + 3] - mean) * rstd * w.w + bb.w;
        }
        write_a(v);
        submit(S_FC1A);
      }
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        await();                                      // fc1 quarters 2 half, 2 half + 1 are in C_W (and, for half 1, fc2 half 0 is in x)
#pragma unroll 1
        for (int g2 = 0; g2 < 2; ++g2) {              // my two 32-column groups of the 128-column half: 2 g2 + ch
          const int grp = 2 * g2 + ch;
          const uint32_t taddr = tmem + lane_addr + C_W + grp * 32;
          ld32f(taddr, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = prm4(P_FC1_B, half * 128 + grp * 32 + 4 * i);
            v[4 * i] = gelu_erf_fast(v[4 * i] + bb.x); v[4 * i + 1] = gelu_erf_fast(v[4 * i + 1] + bb.y);
            v[4 * i + 2] = gelu_erf_fast(v[4 * i + 2] + bb.z); v[4 * i + 3] = gelu_erf_fast(v[4 * i + 3] + bb.w);
          }
          uint32_t hi[16], lo[16];
          split32(v, hi, lo);
          tmem_st16(taddr, hi);                       // in place: [hi 16 columns | lo 16 columns]
          tmem_st16(taddr + 16, lo);
        }
        submit(half == 0 ? S_MID : S_FC2B);
      }
      await();                                        // fc2 half 1 accumulated: x is complete up to the pending biases
      // ---- unbiased-std LayerNorm -> x3 -> q | k | v ----
      {
        ld32f(t_x, x);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = prm4(P_FC2_B, c0 + 4 * i);
          x[4 * i] += t.x; x[4 * i + 1] += t.y; x[4 * i + 2] += t.z; x[4 * i + 3] += t.w;
        }
        float mean, m2;
        stats(x, mean, m2);
        const float rden = 1.0f / (sqrtf(m2 * (1.0f / (E - 1))) + 1e-6f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 w = prm4(P_CLN_A, c0 + 4 * i), bb = prm4(P_CLN_B, c0 + 4 * i);
          x[4 * i] = w.x * (x[4 * i] - mean) * rden + bb.x; x[4 * i + 1] = w.y * (x[4 * i + 1] - mean) * rden + bb.y;
          x[4 * i + 2] = w.z * (x[4 * i + 2] - mean) * rden + bb.z; x[4 * i + 3] = w.w * (x[4 * i + 3] - mean) * rden + bb.w;
        }
        write_a(x);
        submit(S_QKV);
        if (valid) {
          float4* dst = reinterpret_cast<float4*>(p.x3_out + grow * E + c0);
#pragma unroll
          for (int i = 0; i < 8; ++i) dst[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
        }
      }
      await();
#pragma unroll 1
      for (int t3 = 0; t3 < 3; ++t3) {
        ld32f(tmem + lane_addr + t3 * 64 + c0, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bb = prm4(P_QKV_B, t3 * E + c0 + 4 * i);
          v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
        }
        if (valid) {
          if (p.qkv_out) {
            float4* dst = reinterpret_cast<float4*>(p.qkv_out + grow * 3 * E + t3 * E + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          if (p.img_out) {     // fp16 operand image of head `ch`: offset(row, d) = (row/8)*512 + (d/8)*128 + (row%8)*16
            uint8_t* dst = p.img_out + ((size_t)(b * 2 + ch) * 3 + t3) * QKV_IMG + (vert >> 3) * 512 + (vert & 7) * 16;
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(dst + c * 128) = make_uint4(pack_f16(v[8 * c], v[8 * c + 1]), pack_f16(v[8 * c + 2], v[8 * c + 3]),
                                                                    pack_f16(v[8 * c + 4], v[8 * c + 5]), pack_f16(v[8 * c + 6], v[8 * c + 7]));
            if (vert == V - 1) {                      // zero pad row 431
#pragma unroll
              for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(dst + 16 + c * 128) = make_uint4(0, 0, 0, 0);
            }
          }
        }
      }
      // the next tile's x / A writes touch only this thread's own lane and columns, which it has finished reading
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

constexpr int smem_bytes(int J) { return SLOTS * UNIT_BYTES + 2 * J * 128 * 4; }

}  // namespace

// x_in / att_in / kv / outputs as in Chain2Params; prm = 11 device pointers (so_b of the PREVIOUS layer first).
int launch_mdr_chain2(const float* x_in, const float* att_in, const float* kv, const void* blob, const float* const* prm,
                      float* x3_out, float* qkv_out, void* img_out, float* hd_out, int nb, int J, cudaStream_t stream,
                      const ChainEmbed* embed) {
  GATOR_REQUIRE(x_in || (embed && !att_in && !hd_out), "mdr_chain2: x_in may only be null for layer 0 with the embedding operands given");
  static DeviceOnce attr_once;
  static int num_sms[64];
  int dev = 0;
  cudaGetDevice(&dev);
  GATOR_TRY(attr_once.run("mdr_chain2", [&](int d) -> cudaError_t {
    GATOR_CUDA_OK(cudaFuncSetAttribute(mdr_chain2_kernel<17>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(MAXJ)));
    GATOR_CUDA_OK(cudaFuncSetAttribute(mdr_chain2_kernel<19>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(MAXJ)));
    GATOR_CUDA_OK(cudaFuncSetAttribute(mdr_chain2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(MAXJ)));
    return cudaDeviceGetAttribute(&num_sms[d & 63], cudaDevAttrMultiProcessorCount, d);
  }));
  Chain2Params p;
  p.x_in = x_in; p.att_in = att_in; p.kv = kv; p.blob = static_cast<const uint8_t*>(blob);
  for (int i = 0; i < 11; ++i) p.prm[i] = prm[i];
  p.x3_out = x3_out; p.qkv_out = qkv_out; p.img_out = static_cast<uint8_t*>(img_out); p.hd_out = hd_out; p.J = J;
  p.emb = embed ? *embed : ChainEmbed{};
  p.rows_total = (long long)nb * V;
  p.ntiles = (int)((p.rows_total + 127) / 128);
  const int sms = num_sms[dev & 63] > 0 ? num_sms[dev & 63] : 148;
  const int grid = p.ntiles < 2 * sms ? p.ntiles : 2 * sms;
  if (J == 17) mdr_chain2_kernel<17><<<grid, NT, smem_bytes(J), stream>>>(p);
  else if (J == 19) mdr_chain2_kernel<19><<<grid, NT, smem_bytes(J), stream>>>(p);
  else mdr_chain2_kernel<0><<<grid, NT, smem_bytes(J), stream>>>(p);
  return check_launch("mdr_chain2");
}

}  // namespace gator
