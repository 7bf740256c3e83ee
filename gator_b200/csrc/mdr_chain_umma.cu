// Fused MDR layer "chain" (lib/models/MDR.py:140-153 minus the 431x431 softmax core):
// everything that is row-wise between two self-attention cores runs in ONE kernel with the token rows
// resident on the SM:
//
//   x   = x3_prev + selfatt_prev.linears[3](att_prev) + b          (vanilla_transformer_encoder.py:94)   [layers 1,2]
//   q   = LayerNorm1(x) Wq^T                                        (MDR.py:65, :37)
//   a   = softmax_J(q k^T / sqrt(32)) v     per head, k/v of the sample's J joints in shared memory  (:40-43)
//   x  += a Wproj^T + b                                             (:44, :66)
//   x  += fc2(GELU(fc1(LayerNorm2(x))))                             (:68)
//   x3  = a2 (x - mean) / (std_unbiased + 1e-6) + b2                (vanilla_transformer_encoder.py:31-34)
//   qkv = x3 [Wq;Wk;Wv]^T + b                                       (:88-90)          -> x3, qkv to HBM
//
// One CTA = 128 consecutive rows of the flat (sample, vertex) row space, 256 threads, two threads per row (one
// 32-column half each): the fp32 residual row x lives in those threads' registers, every GEMM is a sequence of 64x64 "units"
// (A: 128 x 64 bf16 hi/lo image written by the row owners, W: 64 x 64 bf16 hi/lo image streamed from L2 with
// cp.async into a 3-slot ring, D: 64 TMEM columns per tile), LayerNorm / GELU / the cross-attention run on the
// rows straight out of TMEM.  14 units per layer; per row only x3_prev + att_prev are read and x3 + qkv
// written (1.5 KB instead of 8.7 KB for the kernel-per-op pipeline).
#include "common.cuh"
#include "umma.cuh"

namespace gator {
namespace {

using namespace umma;

constexpr int V = GATOR_V_COARSE;   // 431
constexpr int E = 64;
constexpr int DK = 32;
constexpr int MAXJ = 32;
constexpr int UNIT_IMG = 64 * 64 * 2;          // 8 KB: one 64x64 bf16 image
constexpr int UNIT_BYTES = 2 * UNIT_IMG;       // hi | lo
constexpr int A_IMG = 128 * 64 * 2;            // 16 KB: 128 rows x 64 k
constexpr int A_BUF = 2 * A_IMG;               // hi | lo
constexpr int NUNITS = 14;
enum { U_SO = 0, U_Q = 1, U_PROJ = 2, U_FC1 = 3, U_FC2 = 7, U_QKV = 11 };

// shared memory map (TILES = M=128 tiles per CTA): [tile][buf 2][A_BUF] | W ring [2][UNIT_BYTES] | K|V [J][128] fp32
// TILES = 1: 64 KB + 32 KB + J*512 B (= 105.5 KB for J = 19) -> two CTAs per SM overlap each other's MMA waits
constexpr int W_SLOTS = 2;   // unit u+1 is prefetched into the slot of unit u-1, whose MMAs every thread has waited for
constexpr int smem_bytes(int tiles, int J) { return tiles * 2 * A_BUF + W_SLOTS * UNIT_BYTES + J * 128 * 4; }
// parameter arrays (index into ChainParams::prm)
enum { P_SO_B = 0, P_N1W, P_N1B, P_PROJ_B, P_N2W, P_N2B, P_FC1_B, P_FC2_B, P_CLN_A, P_CLN_B, P_QKV_B };

struct ChainParams {
  const float* x_in;      // (nb*431, 64): layer 0: embedded vertices; layers 1,2: x3 of the previous layer
  const float* att_in;    // (nb*431, 64) self-attention output of the previous layer, or null (layer 0)
  const float* kv;        // (nb*J, 128) this layer's cross-attention K | V
  const uint8_t* blob;    // NUNITS x UNIT_BYTES packed bf16 weight images (unit 0 = previous layer's linears[3])
  const float* prm[11];   // so_b, n1w, n1b, proj_b, n2w, n2b, fc1_b, fc2_b, cln_a, cln_b, qkv_b
  float* x3_out;          // (nb*431, 64)
  float* qkv_out;         // (nb*431, 192)
  float* hd_out;          // non-null: FINAL pass - only x = x_in + att_in Wo^T + b, then hd = x W_head^T + b_head (nb*431, 28);
                          // blob = [linears[3] of the last layer, head (28 rows zero-padded to 64)], prm[0] = so_b, prm[1] = head_b
  int J;
  long long rows_total;   // nb * 431
  int split;              // 1: 3-term bf16 split products
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint4 pack8f(const float* v) {
  return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}
__device__ __forceinline__ uint4 pack8f_res(const float* v, const uint4& hi) {
  return make_uint4(pack_bf16(v[0] - bf16_lo_f(hi.x), v[1] - bf16_hi_f(hi.x)), pack_bf16(v[2] - bf16_lo_f(hi.y), v[3] - bf16_hi_f(hi.y)),
                    pack_bf16(v[4] - bf16_lo_f(hi.z), v[5] - bf16_hi_f(hi.z)), pack_bf16(v[6] - bf16_lo_f(hi.w), v[7] - bf16_hi_f(hi.w)));
}

// row owner writes 8*NCH consecutive k-values of its row into an A image pair (hi, lo), starting at chunk kc0
template <int NCH>
__device__ __forceinline__ void write_a(uint8_t* abuf, int row, int kc0, const float* v) {
  uint8_t* base = abuf + (row >> 3) * 1024 + (row & 7) * 16;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const uint4 hi = pack8f(v + 8 * c);
    *reinterpret_cast<uint4*>(base + (kc0 + c) * 128) = hi;
    *reinterpret_cast<uint4*>(base + A_IMG + (kc0 + c) * 128) = pack8f_res(v + 8 * c, hi);
  }
}


__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// LayerNorm statistics of a 64-wide row held as two 32-wide halves by two threads (different warps): each half does
// its own two-pass mean / M2 and the halves are combined with the parallel-variance formula (see `stats` below).
// TILES x 8 warps: (tile) x (column half 2) x (lane quarter 4)
// JT = compile-time joint count (17 / 19: the cross-attention loops unroll exactly) or 0 = runtime J <= 32
template <int TILES, int JT>
__global__ void __launch_bounds__(256 * TILES, 2 / TILES)
mdr_chain_kernel(ChainParams p) {
  constexpr int NT = 256 * TILES;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float2 xch_buf[2][TILES][128][2];      // [LayerNorm # parity][tile][row][{mine, partner} by column half]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // Tiles are cut from the flat (sample, vertex) row space, so a tile may hold the tail of one sample and the head of the
  // next (431 = 3 x 128 + 47: per-sample tiling would leave every 4th tile 63 % empty).  Everything here is row-wise;
  // the only per-sample data, the cross-attention K|V of the row's sample, is read through L1 (see below).
  const int tile = warp >> 3;                      // which M=128 tile of the CTA
  const int ch = (warp >> 2) & 1;                  // which 32-column half of the 64-wide row this thread owns
  const int row = (warp & 3) * 32 + lane;          // row in tile = TMEM lane
  const long long flat = ((long long)blockIdx.x * TILES + tile) * 128 + row;   // row in the (nb * 431) row space
  const bool valid = flat < p.rows_total;
  const size_t grow = valid ? (size_t)flat : 0;
  const int b = (int)(grow / V);                   // sample of this row
  const int pair_id = 1 + tile * 4 + (warp & 3);   // named barrier shared by the two warps that own the same rows
  constexpr int OFF_W = TILES * 2 * A_BUF;
  uint8_t* a0 = smem + tile * 2 * A_BUF;           // this tile's buffer 0 (n / generic) ...
  uint8_t* a1 = a0 + A_BUF;                        // ... and buffer 1 (GELU(fc1) quarter)
  // K|V (J x 128 fp32 = 9.7 KB per sample): the sample of the tile's first row is staged in shared memory; the rows of
  // a tile that belong to the next sample (30 % of the tiles contain a boundary) read theirs through L1 instead
  // (measured: reading all K|V through L1 costs 28 % per tile, a second shared-memory copy does not fit twice per SM)
  constexpr int OFF_KV = OFF_W + W_SLOTS * UNIT_BYTES;
  float* skv = reinterpret_cast<float*>(smem + OFF_KV);
  const int b_first = (int)(((long long)blockIdx.x * TILES * 128) / V);
  const bool kv_smem = b == b_first;
  const float* kvp = p.kv + (size_t)b * (JT ? JT : p.J) * 128;
  const int J = JT ? JT : p.J;
  constexpr int JU = JT ? JT : MAXJ;               // unrolled trip count of the per-joint loops
  // per-channel parameters come straight from global memory (L1-resident: ~4 KB per layer, shared by all CTAs)
  auto prm32 = [&](int which, int off, float* dst) {     // 32 consecutive parameters as 8 x 16-byte loads
    const float4* src = reinterpret_cast<const float4*>(p.prm[which] + off);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 t = __ldg(src + i);
      dst[4 * i] = t.x; dst[4 * i + 1] = t.y; dst[4 * i + 2] = t.z; dst[4 * i + 3] = t.w;
    }
  };
  const int c0 = ch * 32;                          // first column of this thread's half

  auto prefetch_w = [&](int unit, int slot) {
    const uint8_t* src = p.blob + (size_t)unit * UNIT_BYTES;
    const uint32_t dst = smem_u32(smem + OFF_W + slot * UNIT_BYTES);
#pragma unroll
    for (int i = 0; i < UNIT_BYTES / 16 / NT; ++i) cp_async16(dst + (i * NT + tid) * 16, src + (i * NT + tid) * 16);
    cp_async_commit();
  };

  if (warp == 0) tmem_alloc(&tmem_slot, 128 * TILES);
  if (tid == 32) {
    mbar_init(&bar, 1);
    mbar_init_fence();
  }
  const int first_unit = p.att_in ? U_SO : U_Q;   // (the FINAL pass needs att_in: unit 0 = linears[3], unit 1 = head)
  prefetch_w(first_unit, 0);
  {   // K|V of the tile's first sample: asynchronous 16-byte copies, completed by the cp.async wait of the first unit
    const float* src = p.kv + (size_t)b_first * J * 128;
    const uint32_t dst = smem_u32(skv);
    for (int i = tid; i < J * 32; i += NT) cp_async16(dst + i * 16, src + i * 4);
    cp_async_commit();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t acc = tmem + lane_addr + tile * 128 + c0;   // this thread's 32 columns of the unit accumulator
  const uint32_t acc2 = acc + 64;                            // ... and of the fc2 accumulator
  const uint32_t idesc = idesc_bf16(128, 64);
  uint32_t phase = 0;
  int slot = 0;

  // One 64x64 unit for both tiles: D[tile] (+)= A[tile][buf] . W[slot]^T, then prefetch the next unit's weights.
  auto run_unit = [&](int abuf_idx, int dcol, bool accumulate, int next_unit) {
    cp_async_wait_all();        // this unit's weights have landed (this thread's part)
    fence_proxy_async();        // A operand written by the row owners + W copies -> async proxy
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t w0 = smem_u32(smem + OFF_W + slot * UNIT_BYTES);
#pragma unroll
      for (int t = 0; t < TILES; ++t) {
        const uint32_t abase = smem_u32(smem + (t * 2 + abuf_idx) * A_BUF);
        const uint32_t d = tmem + t * 128 + dcol;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ad = smem_desc(abase + ks * 256, 128, 1024), wd = smem_desc(w0 + ks * 256, 128, 1024);
          const uint32_t accf = (accumulate || ks > 0) ? 1u : 0u;
          if (p.split) {
            mma_bf16(d, smem_desc(abase + A_IMG + ks * 256, 128, 1024), wd, idesc, accf);
            mma_bf16(d, ad, smem_desc(w0 + UNIT_IMG + ks * 256, 128, 1024), idesc, 1);
            mma_bf16(d, ad, wd, idesc, 1);
          } else {
            mma_bf16(d, ad, wd, idesc, accf);
          }
        }
      }
      mma_commit(&bar);
    }
    slot = (slot + 1) % W_SLOTS;
    if (next_unit >= 0) prefetch_w(next_unit, slot);   // that slot was read by the previous unit, whose MMAs were waited for
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
  };
  auto ld32 = [&](uint32_t taddr, float* v) { tmem_ld32(taddr, v); tmem_ld_wait(); };
  auto load_row32 = [&](const float* base, float* v) {
    const float4* src = reinterpret_cast<const float4*>(base + grow * E + c0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 t = valid ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  };
  auto stats = [&](const float* xr, int which, float& mean, float& m2) {
    float2* base = &xch_buf[which & 1][tile][row][0];   // slot reuse is separated by full-CTA barriers
    float m4[4] = {0.f, 0.f, 0.f, 0.f};     // 4 independent chains (a single serial chain costs 4 cycles per add)
#pragma unroll
    for (int i = 0; i < 32; ++i) m4[i & 3] += xr[i];
    const float m = ((m4[0] + m4[1]) + (m4[2] + m4[3])) * (1.0f / 32);
    float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 32; ++i) { const float d = xr[i] - m; q4[i & 3] = fmaf(d, d, q4[i & 3]); }
    const float q = (q4[0] + q4[1]) + (q4[2] + q4[3]);
    base[ch] = make_float2(m, q);
    pair_sync(pair_id);
    const float2 o = base[1 - ch];
    const float dm = m - o.x;
    mean = 0.5f * (m + o.x);
    m2 = (q + o.y) + dm * dm * 16.0f;
  };

  // ---- residual row (this thread's 32 columns) ----
  float x[32], v[32], pw[32], pb[32];
  load_row32(p.x_in, x);
  if (p.att_in) {   // x = x3_prev + att_prev Wo^T + b
    load_row32(p.att_in, v);
    write_a<4>(a0, row, ch * 4, v);
    run_unit(0, 0, false, U_Q);
    ld32(acc, v);
    prm32(P_SO_B, c0, pw);
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] += v[i] + pw[i];
  }
  if (p.hd_out) {
    // final pass after the last self-attention: x = x3 + linears[3](att) + b is complete; apply the MDR head
    // projection [motion_linear | bias_linear | scale_linear] (MDR.py:156-162) and stop
    write_a<4>(a0, row, ch * 4, x);
    run_unit(0, 0, false, -1);
    if (ch == 0) {
      ld32(acc, v);
      if (valid) {
        float* dst = p.hd_out + grow * 28;
#pragma unroll
        for (int i = 0; i < 28; i += 4)
          *reinterpret_cast<float4*>(dst + i) = make_float4(v[i] + __ldg(p.prm[P_N1W] + i), v[i + 1] + __ldg(p.prm[P_N1W] + i + 1),
                                                            v[i + 2] + __ldg(p.prm[P_N1W] + i + 2), v[i + 3] + __ldg(p.prm[P_N1W] + i + 3));
      }
    }
  } else {
  // ---- LayerNorm1 -> q ----
  {
    float mean, m2;
    stats(x, 0, mean, m2);
    const float rstd = rsqrtf(m2 * (1.0f / E) + 1e-5f);
    prm32(P_N1W, c0, pw);
    prm32(P_N1B, c0, pb);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (x[i] - mean) * rstd * pw[i] + pb[i];
    write_a<4>(a0, row, ch * 4, v);
  }
  run_unit(0, 0, false, U_PROJ);
  // ---- cross attention against the sample's J joints: this thread owns head `ch` of its row ----
  {
    float q[DK];
    ld32(acc, q);
    auto cross = [&](const float* kvb, auto ld4) {
    float s[JU];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < JU; ++j) {
      if (JT || j < J) {
        const float4* kr = reinterpret_cast<const float4*>(kvb + j * 128 + c0);
        float a0_ = 0.f, a1_ = 0.f, a2_ = 0.f, a3_ = 0.f;       // 4 independent FMA chains
#pragma unroll
        for (int d4 = 0; d4 < DK / 4; ++d4) {
          const float4 kk = ld4(kr + d4);
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
          const float4 vv = ld4(vr + d4);
          v[4 * d4] = fmaf(pj, vv.x, v[4 * d4]); v[4 * d4 + 1] = fmaf(pj, vv.y, v[4 * d4 + 1]);
          v[4 * d4 + 2] = fmaf(pj, vv.z, v[4 * d4 + 2]); v[4 * d4 + 3] = fmaf(pj, vv.w, v[4 * d4 + 3]);
        }
      }
    }
    };
    if (kv_smem) cross(skv, [](const float4* q_) { return *q_; });
    else cross(kvp, [](const float4* q_) { return __ldg(q_); });
    write_a<4>(a0, row, ch * 4, v);
  }
  run_unit(0, 0, false, U_FC1);
  ld32(acc, v);
  prm32(P_PROJ_B, c0, pw);
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] += v[i] + pw[i];
  // ---- LayerNorm2 -> MLP ----
  {
    float mean, m2;
    stats(x, 1, mean, m2);
    const float rstd = rsqrtf(m2 * (1.0f / E) + 1e-5f);
    prm32(P_N2W, c0, pw);
    prm32(P_N2B, c0, pb);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (x[i] - mean) * rstd * pw[i] + pb[i];
    write_a<4>(a0, row, ch * 4, v);
  }
#pragma unroll 1
  for (int qd = 0; qd < 4; ++qd) {
    run_unit(0, 0, false, U_FC2 + qd);                     // fc1 quarter qd from LN2(x) (buffer 0 stays intact)
    ld32(acc, v);
    prm32(P_FC1_B, qd * 64 + c0, pw);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf_fast(v[i] + pw[i]);
    write_a<4>(a1, row, ch * 4, v);
    run_unit(1, 64, qd > 0, qd < 3 ? U_FC1 + qd + 1 : U_QKV);   // fc2 += GELU(.) W2[:, quarter]
  }
  ld32(acc2, v);
  prm32(P_FC2_B, c0, pw);
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] += v[i] + pw[i];
  // ---- unbiased-std LayerNorm -> x3 ----
  {
    float mean, m2;
    stats(x, 2, mean, m2);
    const float denom = sqrtf(m2 * (1.0f / (E - 1))) + 1e-6f;
    prm32(P_CLN_A, c0, pw);
    prm32(P_CLN_B, c0, pb);
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = pw[i] * (x[i] - mean) / denom + pb[i];
    write_a<4>(a0, row, ch * 4, x);
    if (valid) {
      float4* dst = reinterpret_cast<float4*>(p.x3_out + grow * E + c0);
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
    }
  }
  // ---- q | k | v projections of the self-attention that follows ----
#pragma unroll 1
  for (int t3 = 0; t3 < 3; ++t3) {
    run_unit(0, 0, false, t3 < 2 ? U_QKV + t3 + 1 : -1);
    ld32(acc, v);
    if (valid) {
      float4* dst = reinterpret_cast<float4*>(p.qkv_out + grow * 3 * E + t3 * E + c0);
      const float* bq = p.prm[P_QKV_B] + t3 * E + c0;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        dst[i] = make_float4(v[4 * i] + __ldg(bq + 4 * i), v[4 * i + 1] + __ldg(bq + 4 * i + 1), v[4 * i + 2] + __ldg(bq + 4 * i + 2), v[4 * i + 3] + __ldg(bq + 4 * i + 3));
    }
  }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128 * TILES);
}

}  // namespace

// x_in / att_in / kv / outputs as in ChainParams; prm = 11 device pointers (so_b of the PREVIOUS layer first).
int launch_mdr_chain(const float* x_in, const float* att_in, const float* kv, const void* blob, const float* const* prm,
                     float* x3_out, float* qkv_out, float* hd_out, int nb, int J, bool split, cudaStream_t stream) {
  static DeviceOnce attr_once;
  GATOR_TRY(attr_once.run("mdr_chain", [&](int) -> cudaError_t {
    GATOR_CUDA_OK(cudaFuncSetAttribute(mdr_chain_kernel<1, 17>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(1, MAXJ)));
    GATOR_CUDA_OK(cudaFuncSetAttribute(mdr_chain_kernel<1, 19>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(1, MAXJ)));
    return cudaFuncSetAttribute(mdr_chain_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(1, MAXJ));
  }));
  ChainParams p;
  p.x_in = x_in; p.att_in = att_in; p.kv = kv; p.blob = static_cast<const uint8_t*>(blob);
  for (int i = 0; i < 11; ++i) p.prm[i] = prm[i];
  p.x3_out = x3_out; p.qkv_out = qkv_out; p.hd_out = hd_out; p.J = J; p.split = split ? 1 : 0;
  p.rows_total = (long long)nb * V;
  // one 128-row tile per CTA; two CTAs per SM as long as J <= 24 (2 x ~107 KB)
  const unsigned grid = (unsigned)((p.rows_total + 127) / 128);
  if (J == 17) mdr_chain_kernel<1, 17><<<grid, 256, smem_bytes(1, J), stream>>>(p);
  else if (J == 19) mdr_chain_kernel<1, 19><<<grid, 256, smem_bytes(1, J), stream>>>(p);
  else mdr_chain_kernel<1, 0><<<grid, 256, smem_bytes(1, J), stream>>>(p);
  return check_launch("mdr_chain");
}

}  // namespace gator
