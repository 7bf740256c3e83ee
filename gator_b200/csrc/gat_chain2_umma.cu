// Fused GAT blocks (lib/models/GAT.py:33-43 x depth) on tcgen05, round-2 kernel: ONE launch runs all GATBlocks of the
// lifter for a tile of whole samples.
//
//   n  = LayerNorm1(x)
//   a  = proj(softmax_J(q k^T / 4 + hop_path_bias) v)        8 heads of 16, per sample        (modules.py:121-138)
//   g  = (A o I)(M o n W0) + (A o (1-I))(M o n W1) + b       modulated graph convolution      (modules.py:243-255)
//   x += linearback([1[hop<=1] L0(a+g) | 1[hop==2] L1(a+g)])                                  (modules.py:158-177)
//   x += fc2(GELU(fc1(LayerNorm2(x))))                                                        (modules.py:188-196)
//
// CTA = 128 token rows = S whole samples (S = 128 / J: 6 for J = 19, 7 for J = 17), 512 threads = 4 threads per row
// (32 of the 128 channels each).  Against the round-1 kernel (36 block-wide MMA round trips per GATBlock, A operands
// staged through 128 KB of shared memory, the attention math on a quarter of the threads):
//  * every A operand and the residual stream x live in TENSOR MEMORY (x is the accumulator of linearback and fc2; A
//    operands are written by the row owners with tcgen05.st and read with tcgen05.mma [d], [a], b-desc; GELU(fc1) is
//    converted in place and read back as the A operand of fc2) - 512 columns: x 128 | A 128 | work 256;
//  * 9 MMA round trips per GATBlock: {q, k|v heads 0-3}, {k|v heads 4-7}, proj, {h0, h1}, {L0|L1}, linearback,
//    fc1 half 0, {fc2 half 0, fc1 half 1}, fc2 half 1 - 34 weight pieces of 32 KB stream through a 4-slot TMA ring
//    (cp.async.bulk + mbarriers), requested several pieces ahead by the leader lane, which also issues the MMAs;
//  * the attention core runs on all 512 threads (thread = (row, head) for four heads at a time, K|V of those heads
//    exchanged through 64 KB of shared memory) while the MMAs of the other four heads' K|V are in flight;
//  * the J x J mixes (graph convolution, hop masks) go through the same shared-memory buffer in fp32.
#include "common.cuh"
#include "umma.cuh"

namespace gator {
namespace {

using namespace umma;

constexpr int C = 128;
constexpr int MAXJ = 32;
constexpr int NT = 512;
constexpr int PIECE_IMG = 64 * 128 * 2;        // 16 KB: one bf16 image of a piece (64x128 or 128x64)
constexpr int PIECE_BYTES = 2 * PIECE_IMG;     // hi | lo
constexpr int PIECES = 34;                     // per block, in consumption order (see GAT.py: pack)
constexpr int SLOTS = 4;
constexpr int XS = 132;                        // row stride (floats) of the fp32 exchange buffer: rows fall in distinct banks
constexpr int OFF_RING = 0;
constexpr int OFF_XCH = SLOTS * PIECE_BYTES;               // 131072: fp32 exchange [128][132]            67 584
constexpr int OFF_BIAS = OFF_XCH + 128 * XS * 4;           // 198656: attention bias [8][J][J] fp32   <= 11 552 (J = 19)
constexpr int smem_bytes(int J) { return OFF_BIAS + 8 * J * J * 4; }
// tensor-memory columns
constexpr int C_X = 0;        // residual stream (fp32, 128 columns)
constexpr int C_A = 128;      // A operand of K = 128: hi columns [128,192), lo columns [192,256)
constexpr int C_W = 256;      // work region (256 columns)

struct GatChain2Params {
  float* x;                       // (rows, 128) in/out
  int rows;                       // B * J
  int J, S, depth;
  const uint8_t* const* blobs;    // DEVICE array [depth] of piece blobs (34 x 32 KB each)
  const float* const* prm;        // DEVICE array [depth * 14] of per-block fp32 parameter arrays (see PRM_*)
  const float* attn_bias;         // (8, J, J)
  const float* mask1;             // (J, J) 1[hop <= 1]
  const float* mask2;             // (J, J) 1[hop == 2]
};
enum { PRM_LN1W = 0, PRM_LN1B, PRM_QKVB, PRM_PROJB, PRM_GCNM, PRM_ADIAG, PRM_AOFF, PRM_GCNB, PRM_XFB01, PRM_XFBB,
       PRM_LN2W, PRM_LN2B, PRM_FC1B, PRM_FC2B, PRM_COUNT };
// MMA groups of a block, in order
enum Step { S_QKV0, S_KV1, S_PROJ, S_GCN, S_XF01, S_XFB, S_FC1A, S_MID, S_FC2B };

struct Bars {
  uint64_t w_full[SLOTS], w_empty[SLOTS];
  uint64_t a_ready, d_ready;
};

__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// n fp32 values -> n/2 packed bf16x2 "hi" registers and n/2 packed residual registers
template <int N>
__device__ __forceinline__ void split(const float* v, uint32_t* hi, uint32_t* lo) {
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    const uint32_t h = pack_bf16(v[2 * i], v[2 * i + 1]);
    hi[i] = h;
    lo[i] = pack_bf16(v[2 * i] - bf16_lo_f(h), v[2 * i + 1] - bf16_hi_f(h));
  }
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

template <int JT>
__global__ void __launch_bounds__(NT, 1) gat_chain2_kernel(GatChain2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ Bars bars;
  __shared__ uint32_t tmem_slot;
  __shared__ float2 xch_ln[2][128][4];              // LayerNorm partials [parity][row][column quarter]
  __shared__ uint32_t hopbits[2][MAXJ];             // hop masks as bit sets per query joint
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int J = JT ? JT : p.J;
  constexpr int JU = JT ? JT : MAXJ;
  const int S = p.S;
  const int cq = warp >> 2;                         // column quarter: channels [32cq, 32cq+32)
  const int row = (warp & 3) * 32 + lane;           // row in tile = tensor-memory lane
  const int row0 = blockIdx.x * S * J;              // first global row of the tile
  const int nrows = min(S * J, p.rows - row0);      // valid rows (whole samples)
  const bool valid = row < nrows;
  const int samp = valid ? row / J : 0;             // sample within the tile
  const int ji = valid ? row - samp * J : 0;        // joint index
  const int srow0 = samp * J;                       // first row of this row's sample
  float* xch = reinterpret_cast<float*>(smem + OFF_XCH);
  float* sbias = reinterpret_cast<float*>(smem + OFF_BIAS);
  const bool leader = warp == 0;                    // warp 0 (all lanes, one elected to issue) is also TMA producer and MMA issuer

  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 32) {
    for (int s = 0; s < SLOTS; ++s) { mbar_init(&bars.w_full[s], 1); mbar_init(&bars.w_empty[s], 1); }
    mbar_init(&bars.a_ready, NT / 32);
    mbar_init(&bars.d_ready, 1);
    mbar_init_fence();
  }
  for (int i = tid; i < 8 * J * J; i += NT) sbias[i] = p.attn_bias[i];
  if (tid < 2 * J) {
    const int which = tid / J, i = tid - which * J;
    const float* m = which ? p.mask2 : p.mask1;
    uint32_t bits = 0;
    for (int j = 0; j < J; ++j) bits |= (m[i * J + j] != 0.f ? 1u : 0u) << j;
    hopbits[which][i] = bits;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t t_lane = tmem + lane_addr;

  // ---- leader warp: TMA producer + MMA issuer (executed by the whole warp, instructions issued by one elected lane) ----
  const int total_units = p.depth * PIECES;
  int u_use = 0, u_load = 0;
  uint32_t ph_a = 0;
  // Request pieces up to SLOTS ahead.  A slot is reused once the MMAs that read it are done (w_empty); the leader only
  // blocks on that when the very next piece it has to issue has not been requested yet, otherwise it probes and moves on.
  auto refill = [&]() {
    while (u_load < total_units && u_load - u_use < SLOTS) {
      const int s = u_load % SLOTS;
      if (u_load >= SLOTS) {
        const uint32_t par = ((u_load / SLOTS) - 1) & 1;
        if (!mbar_test(&bars.w_empty[s], par)) {
          if (u_load > u_use) break;
          mbar_wait(&bars.w_empty[s], par);
        }
      }
      const uint8_t* src = p.blobs[u_load / PIECES] + (size_t)(u_load % PIECES) * PIECE_BYTES;
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars.w_full[s], PIECE_BYTES);
        bulk_copy_g2s(smem + OFF_RING + s * PIECE_BYTES, src, PIECE_BYTES, &bars.w_full[s]);
      }
      ++u_load;
    }
  };
  // One piece.  narrow: D[dcol, +64) = A(K = 128: 8 k-steps) . W(64 x 128)^T.   wide: D[dcol, +128) (+)= A(K = 64: 4 k-steps) . W(128 x 64)^T.
  // A k-step ks: hi columns a_hi + (ks >> 1) * a_grp2 + (ks & 1) * 8, lo a_lo_off further.
  auto piece = [&](bool wide, uint32_t dcol, bool accumulate, uint32_t a_hi, uint32_t a_lo_off, uint32_t a_grp2) {
    refill();
    const int s = u_use % SLOTS;
    mbar_wait(&bars.w_full[s], (u_use / SLOTS) & 1);
    const uint32_t w0 = smem_u32(smem + OFF_RING + s * PIECE_BYTES);
    const uint32_t idesc = wide ? idesc_bf16(128, 128) : idesc_bf16(128, 64);
    const uint32_t w_sbo = wide ? 1024u : 2048u;
    const int ksteps = wide ? 4 : 8;
    if (elect_one()) {
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint32_t ah = tmem + a_hi + (ks >> 1) * a_grp2 + (ks & 1) * 8, al = ah + a_lo_off;
        const uint64_t wh = smem_desc(w0 + ks * 256, 128, w_sbo), wl = smem_desc(w0 + PIECE_IMG + ks * 256, 128, w_sbo);
        mma_ts(tmem + dcol, al, wh, idesc, (accumulate || ks > 0) ? 1u : 0u);
        mma_ts(tmem + dcol, ah, wl, idesc, 1);
        mma_ts(tmem + dcol, ah, wh, idesc, 1);
      }
      mma_commit(&bars.w_empty[s]);
    }
    __syncwarp();
    ++u_use;
    refill();
  };
  // A region (K = 128): k-step ks at C_A + 8 ks, lo 64 columns further; a K = 64 slice kh starts at k-step 4 kh
  auto narrow_from_a = [&](uint32_t dcol) { piece(false, dcol, false, C_A, 64, 16); };
  auto wide_from_a = [&](uint32_t dcol, int kh, bool acc) { piece(true, dcol, acc, C_A + kh * 32, 64, 16); };
  // GELU'd fc1 half in C_W: 32-column groups [hi 16 | lo 16]; fc2 slice q4 (0..3 within the half) = groups 2 q4, 2 q4 + 1
  auto wide_from_w = [&](int q4) { piece(true, C_X, true, C_W + q4 * 64, 16, 32); };
  auto issue = [&](int step) {                       // leader only
    mbar_wait(&bars.a_ready, ph_a);
    ph_a ^= 1;
    tc_fence_after();
    switch (step) {
      case S_QKV0: for (int u = 0; u < 4; ++u) narrow_from_a(C_W + 64 * u); break;          // q (2 units) | k heads 0-3 | v heads 0-3
      case S_KV1: narrow_from_a(C_W + 128); narrow_from_a(C_W + 192); break;                 // k heads 4-7 | v heads 4-7
      case S_PROJ: wide_from_a(C_W, 0, false); wide_from_a(C_W, 1, true); break;             // a = o Wp^T
      case S_GCN: wide_from_a(C_W, 0, false); wide_from_a(C_W, 1, true);                     // h0 = n W0
                  wide_from_a(C_W + 128, 0, false); wide_from_a(C_W + 128, 1, true); break;  // h1 = n W1
      case S_XF01: for (int u = 0; u < 3; ++u) narrow_from_a(C_W + 64 * u); break;           // [L0 | L1](s), 192 columns
      case S_XFB: wide_from_a(C_X, 0, true); wide_from_a(C_X, 1, true);                      // x += linearback(f): K = 128 from the A region ...
                  piece(true, C_X, true, C_W + 192, 32, 16); break;                          // ... + K = 64 (hop-2 part) from work columns [192,256)
      case S_FC1A: for (int u = 0; u < 4; ++u) narrow_from_a(C_W + 64 * u); break;           // fc1 units 0-3
      case S_MID: for (int q4 = 0; q4 < 4; ++q4) wide_from_w(q4);                            // x += fc2 half 0
                  for (int u = 0; u < 4; ++u) narrow_from_a(C_W + 64 * u); break;            // fc1 units 4-7
      default: for (int q4 = 0; q4 < 4; ++q4) wide_from_w(q4); break;                        // S_FC2B: x += fc2 half 1
    }
    if (elect_one()) mma_commit(&bars.d_ready);
    __syncwarp();
  };
  if (leader) refill();

  uint32_t ph_d = 0;
  int ln_count = 0;
  auto submit = [&](int step) {                      // my part of the A operands is written (and my reads of the work region are done)
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
  };
  auto ld32f = [&](uint32_t taddr, float* v) {
    uint32_t r[32];
    tmem_ld32_async(taddr, r);
    tmem_ld_wait_dep32(r);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
  };
  auto ld16f = [&](uint32_t taddr, float* v) {
    uint32_t r[16];
    tmem_ld16_async(taddr, r);
    tmem_ld_wait_dep16(r);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
  };
  auto st32f = [&](uint32_t taddr, const float* v) {
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(v[i]);
    tmem_st32(taddr, r);
  };
  auto write_a32 = [&](const float* v) {             // my 32 k-values [32cq, 32cq+32) of the K = 128 A operand
    uint32_t hi[16], lo[16];
    split<32>(v, hi, lo);
    tmem_st16(t_lane + C_A + cq * 16, hi);
    tmem_st16(t_lane + C_A + 64 + cq * 16, lo);
  };
  auto prm4 = [&](const float* base, int off) { return __ldg(reinterpret_cast<const float4*>(base + off)); };
  auto add_prm32 = [&](float* v, const float* base, int off) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 t = prm4(base, off + 4 * i);
      v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
    }
  };
  // LayerNorm statistics over the 128 channels of a row held by 4 threads (32 each)
  auto stats128 = [&](const float* xr, float& mean, float& rstd) {
    float2* base = &xch_ln[ln_count & 1][row][0];
    ++ln_count;
    float m4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 32; ++i) m4[i & 3] += xr[i];
    const float m = ((m4[0] + m4[1]) + (m4[2] + m4[3])) * (1.0f / 32);
    float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 32; ++i) { const float d = xr[i] - m; q4[i & 3] = fmaf(d, d, q4[i & 3]); }
    base[cq] = make_float2(m, (q4[0] + q4[1]) + (q4[2] + q4[3]));
    group_sync(1 + (warp & 3));
    const float2 a = base[0], b = base[1], c = base[2], d = base[3];
    mean = 0.25f * ((a.x + b.x) + (c.x + d.x));
    const float da = a.x - mean, db = b.x - mean, dc = c.x - mean, dd = d.x - mean;
    const float m2 = ((a.y + b.y) + (c.y + d.y)) + 32.0f * ((da * da + db * db) + (dc * dc + dd * dd));
    rstd = rsqrtf(m2 * (1.0f / C) + 1e-5f);
  };
  auto normalize = [&](const float* xr, float mean, float rstd, const float* w, const float* b, float* out) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 ww = prm4(w, cq * 32 + 4 * i), bb = prm4(b, cq * 32 + 4 * i);
      out[4 * i] = (xr[4 * i] - mean) * rstd * ww.x + bb.x; out[4 * i + 1] = (xr[4 * i + 1] - mean) * rstd * ww.y + bb.y;
      out[4 * i + 2] = (xr[4 * i + 2] - mean) * rstd * ww.z + bb.z; out[4 * i + 3] = (xr[4 * i + 3] - mean) * rstd * ww.w + bb.w;
    }
  };

  const uint32_t t_x = t_lane + C_X + cq * 32;       // my quarter of x
  float x[32], v[32];
  {
    const float4* src = reinterpret_cast<const float4*>(p.x + (size_t)(row0 + (valid ? row : 0)) * C + cq * 32);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 t = valid ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      x[4 * i] = t.x; x[4 * i + 1] = t.y; x[4 * i + 2] = t.z; x[4 * i + 3] = t.w;
    }
    st32f(t_x, x);
    tmem_st_wait();
  }

#pragma unroll 1
  for (int blk = 0; blk < p.depth; ++blk) {
    const float* const* prm = p.prm + blk * PRM_COUNT;
    // ---- LayerNorm1 -> A ----   (x lives in tensor memory; it is re-read wherever it is needed)
    float mean1, rstd1;
    ld32f(t_x, x);
    stats128(x, mean1, rstd1);
    normalize(x, mean1, rstd1, prm[PRM_LN1W], prm[PRM_LN1B], v);
    write_a32(v);
    submit(S_QKV0);
    // ---- attention: four heads at a time, thread = (row, head 4 half + cq) ----
    float o[2][16];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      await();                                        // half 0: q | k03 | v03;  half 1: k47 | v47
      const int h = half * 4 + cq;
      float kk[16], vv[16], q[16];
      ld16f(t_lane + C_W + 128 + cq * 16, kk);        // k of head h, this row
      ld16f(t_lane + C_W + 192 + cq * 16, vv);
      ld16f(t_lane + C_W + h * 16, q);
      {
        const float* bq = prm[PRM_QKVB];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 bk = prm4(bq, 128 + h * 16 + 4 * i), bv = prm4(bq, 256 + h * 16 + 4 * i), b0 = prm4(bq, h * 16 + 4 * i);
          kk[4 * i] += bk.x; kk[4 * i + 1] += bk.y; kk[4 * i + 2] += bk.z; kk[4 * i + 3] += bk.w;
          vv[4 * i] += bv.x; vv[4 * i + 1] += bv.y; vv[4 * i + 2] += bv.z; vv[4 * i + 3] += bv.w;
          q[4 * i] += b0.x; q[4 * i + 1] += b0.y; q[4 * i + 2] += b0.z; q[4 * i + 3] += b0.w;
        }
      }
      if (half == 1) {
        // o of heads 0-3 -> A operand (the MMAs that read n from the A region, k47 | v47, are complete)
        uint32_t hi[8], lo[8];
        split<16>(o[0], hi, lo);
        tmem_st8(t_lane + C_A + cq * 8, hi);
        tmem_st8(t_lane + C_A + 64 + cq * 8, lo);
        __syncthreads();                              // everyone is done reading heads 0-3 from the exchange buffer
      }
      {
        float* dst = xch + row * XS + cq * 32;        // [k16 | v16] of head (half, cq)
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          *reinterpret_cast<float4*>(dst + i) = make_float4(kk[i], kk[i + 1], kk[i + 2], kk[i + 3]);
          *reinterpret_cast<float4*>(dst + 16 + i) = make_float4(vv[i], vv[i + 1], vv[i + 2], vv[i + 3]);
        }
      }
      tc_fence_before();
      __syncthreads();                                // k | v of the four heads staged; the work columns [128,256) are free again
      if (half == 0) submit(S_KV1);                   // k | v of heads 4-7 compute while heads 0-3 are attended to
      {
        float s[JU];
        float mx = -INFINITY;
        const float* bias_row = sbias + (h * J + ji) * J;
        const float* kvb = xch + srow0 * XS + cq * 32;
#pragma unroll
        for (int j = 0; j < JU; ++j) {
          if (JT || j < J) {
            const float4* kr = reinterpret_cast<const float4*>(kvb + j * XS);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int d4 = 0; d4 < 4; ++d4) {
              const float4 k4 = kr[d4];
              a0 = fmaf(q[4 * d4], k4.x, a0); a1 = fmaf(q[4 * d4 + 1], k4.y, a1);
              a2 = fmaf(q[4 * d4 + 2], k4.z, a2); a3 = fmaf(q[4 * d4 + 3], k4.w, a3);
            }
            s[j] = ((a0 + a1) + (a2 + a3)) * 0.25f + bias_row[j];
            mx = fmaxf(mx, s[j]);
          }
        }
        float l = 0.f;
#pragma unroll
        for (int j = 0; j < JU; ++j)
          if (JT || j < J) { s[j] = expf(s[j] - mx); l += s[j]; }
        const float inv = 1.0f / l;
#pragma unroll
        for (int d = 0; d < 16; ++d) o[half][d] = 0.f;
#pragma unroll
        for (int j = 0; j < JU; ++j) {
          if (JT || j < J) {
            const float pj = s[j] * inv;
            const float4* vr = reinterpret_cast<const float4*>(kvb + j * XS + 16);
#pragma unroll
            for (int d4 = 0; d4 < 4; ++d4) {
              const float4 v4 = vr[d4];
              o[half][4 * d4] = fmaf(pj, v4.x, o[half][4 * d4]); o[half][4 * d4 + 1] = fmaf(pj, v4.y, o[half][4 * d4 + 1]);
              o[half][4 * d4 + 2] = fmaf(pj, v4.z, o[half][4 * d4 + 2]); o[half][4 * d4 + 3] = fmaf(pj, v4.w, o[half][4 * d4 + 3]);
            }
          }
        }
      }
    }
    {
      uint32_t hi[8], lo[8];
      split<16>(o[1], hi, lo);
      tmem_st8(t_lane + C_A + 32 + cq * 8, hi);       // heads 4-7: k-values [64 + 16cq, +16)
      tmem_st8(t_lane + C_A + 64 + 32 + cq * 8, lo);
    }
    submit(S_PROJ);
    await();
    // ---- s = a + g, built up in registers ----
    float s_[32];
    ld32f(t_lane + C_W + cq * 32, s_);
    add_prm32(s_, prm[PRM_PROJB], cq * 32);
    ld32f(t_x, x);
    normalize(x, mean1, rstd1, prm[PRM_LN1W], prm[PRM_LN1B], v);   // n again (the A region held o in between)
    write_a32(v);
    submit(S_GCN);
    await();
    {
      float m[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 t = prm4(prm[PRM_GCNM], ji * C + cq * 32 + 4 * i);   // M[ji, channels]
        m[4 * i] = t.x; m[4 * i + 1] = t.y; m[4 * i + 2] = t.z; m[4 * i + 3] = t.w;
      }
      ld32f(t_lane + C_W + cq * 32, v);               // h0
      const float ad = __ldg(prm[PRM_ADIAG] + ji);
#pragma unroll
      for (int i = 0; i < 32; ++i) s_[i] = fmaf(ad, m[i] * v[i], s_[i]);
      add_prm32(s_, prm[PRM_GCNB], cq * 32);
      ld32f(t_lane + C_W + 128 + cq * 32, v);         // h1
      float* dst = xch + row * XS + cq * 32;          // stage M o h1 for the other joints of the sample (attention is done with the buffer)
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(dst + i) = make_float4(m[i] * v[i], m[i + 1] * v[i + 1], m[i + 2] * v[i + 2], m[i + 3] * v[i + 3]);
    }
    __syncthreads();
    {
      const float* aoff = prm[PRM_AOFF] + ji * J;
      const float* sb = xch + srow0 * XS + cq * 32;
#pragma unroll
      for (int j = 0; j < JU; ++j) {
        if (JT || j < J) {
          const float aij = __ldg(aoff + j);
          const float4* sr = reinterpret_cast<const float4*>(sb + j * XS);
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            const float4 t = sr[i4];
            s_[4 * i4] = fmaf(aij, t.x, s_[4 * i4]); s_[4 * i4 + 1] = fmaf(aij, t.y, s_[4 * i4 + 1]);
            s_[4 * i4 + 2] = fmaf(aij, t.z, s_[4 * i4 + 2]); s_[4 * i4 + 3] = fmaf(aij, t.w, s_[4 * i4 + 3]);
          }
        }
      }
    }
    write_a32(s_);
    submit(S_XF01);                                   // (every thread's reads of the exchange buffer precede its arrival)
    await();
    // ---- X_Feat hop mixes: f = [1[hop<=1] @ L0(s) (128) | 1[hop==2] @ L1(s) (16, zero-padded to 64)] ----
    {
      ld32f(t_lane + C_W + cq * 32, v);
      add_prm32(v, prm[PRM_XFB01], cq * 32);
      float* dst = xch + row * XS + cq * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      float y1[16];
      ld16f(t_lane + C_W + 128 + cq * 16, y1);        // L1 outputs 16cq .. 16cq+15 of the 64-wide (16 valid + zero rows) unit
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 t = prm4(prm[PRM_XFB01], 128 + cq * 16 + 4 * i);
        y1[4 * i] += t.x; y1[4 * i + 1] += t.y; y1[4 * i + 2] += t.z; y1[4 * i + 3] += t.w;
      }
      __syncthreads();
      const uint32_t bits1 = hopbits[0][ji], bits2 = hopbits[1][ji];
      const float* sb = xch + srow0 * XS + cq * 32;
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
      for (int j = 0; j < JU; ++j) {
        if ((JT || j < J) && ((bits1 >> j) & 1u)) {
          const float4* sr = reinterpret_cast<const float4*>(sb + j * XS);
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            const float4 t = sr[i4];
            v[4 * i4] += t.x; v[4 * i4 + 1] += t.y; v[4 * i4 + 2] += t.z; v[4 * i4 + 3] += t.w;
          }
        }
      }
      write_a32(v);
      __syncthreads();                                // pass 1 reads done: the buffer takes the 64 hop-2 columns (row stride XS)
      float* d1 = xch + row * XS + cq * 16;
#pragma unroll
      for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(d1 + i) = make_float4(y1[i], y1[i + 1], y1[i + 2], y1[i + 3]);
      __syncthreads();
      const float* sb1 = xch + srow0 * XS + cq * 16;
#pragma unroll
      for (int i = 0; i < 16; ++i) y1[i] = 0.f;
#pragma unroll
      for (int j = 0; j < JU; ++j) {
        if ((JT || j < J) && ((bits2 >> j) & 1u)) {
          const float4* sr = reinterpret_cast<const float4*>(sb1 + j * XS);
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const float4 t = sr[i4];
            y1[4 * i4] += t.x; y1[4 * i4 + 1] += t.y; y1[4 * i4 + 2] += t.z; y1[4 * i4 + 3] += t.w;
          }
        }
      }
      uint32_t hi[8], lo[8];
      split<16>(y1, hi, lo);                          // K = 64 operand in work columns [192,256): hi 32 | lo 32
      tmem_st8(t_lane + C_W + 192 + cq * 8, hi);
      tmem_st8(t_lane + C_W + 192 + 32 + cq * 8, lo);
    }
    submit(S_XFB);
    await();
    // ---- MLP ----
    ld32f(t_x, x);
    add_prm32(x, prm[PRM_XFBB], cq * 32);
    st32f(t_x, x);                                    // fc2 accumulates onto x + linearback bias
    {
      float mean, rstd;
      stats128(x, mean, rstd);
      normalize(x, mean, rstd, prm[PRM_LN2W], prm[PRM_LN2B], v);
      write_a32(v);
    }
    submit(S_FC1A);
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      await();                                        // fc1 units 4 half .. 4 half + 3 are in the work columns
#pragma unroll 1
      for (int g2 = 0; g2 < 2; ++g2) {                // my two 32-column groups: columns [64cq + 32 g2, +32)
        const uint32_t taddr = t_lane + C_W + cq * 64 + g2 * 32;
        ld32f(taddr, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bb = prm4(prm[PRM_FC1B], half * 256 + cq * 64 + g2 * 32 + 4 * i);
          v[4 * i] = gelu_erf_fast(v[4 * i] + bb.x); v[4 * i + 1] = gelu_erf_fast(v[4 * i + 1] + bb.y);
          v[4 * i + 2] = gelu_erf_fast(v[4 * i + 2] + bb.z); v[4 * i + 3] = gelu_erf_fast(v[4 * i + 3] + bb.w);
        }
        uint32_t hi[16], lo[16];
        split<32>(v, hi, lo);
        tmem_st16(taddr, hi);                         // in place: [hi 16 columns | lo 16 columns]
        tmem_st16(taddr + 16, lo);
      }
      submit(half == 0 ? S_MID : S_FC2B);
    }
    await();
    ld32f(t_x, x);
    add_prm32(x, prm[PRM_FC2B], cq * 32);
    if (blk + 1 < p.depth) {                          // the next block's linearback accumulates onto the complete x
      st32f(t_x, x);
      tmem_st_wait();
    }
  }
  if (valid) {
    float4* dst = reinterpret_cast<float4*>(p.x + (size_t)(row0 + row) * C + cq * 32);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

bool gat_chain_supported(int J) { return J >= 2 && J <= 21; }   // attention-bias table [8][J][J] must fit beside the ring

int launch_gat_chain(float* x, int rows, int J, int depth, const void* const* blobs_dev, const float* const* prm_dev,
                     const float* attn_bias, const float* mask1, const float* mask2, cudaStream_t stream) {
  static DeviceOnce attr_once;
  GATOR_TRY(attr_once.run("gat_chain2", [&](int) -> cudaError_t {
    GATOR_CUDA_OK(cudaFuncSetAttribute(gat_chain2_kernel<17>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(17)));
    GATOR_CUDA_OK(cudaFuncSetAttribute(gat_chain2_kernel<19>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(19)));
    return cudaFuncSetAttribute(gat_chain2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(21));
  }));
  GATOR_REQUIRE(gat_chain_supported(J), "gat_chain: num_joint=%d does not fit the fused kernel", J);
  GatChain2Params p;
  p.x = x; p.rows = rows; p.J = J; p.S = 128 / J; p.depth = depth;
  p.blobs = reinterpret_cast<const uint8_t* const*>(blobs_dev);
  p.prm = prm_dev; p.attn_bias = attn_bias; p.mask1 = mask1; p.mask2 = mask2;
  const int rows_per_tile = p.S * J;
  const int tiles = (rows + rows_per_tile - 1) / rows_per_tile;
  if (J == 17) gat_chain2_kernel<17><<<tiles, NT, smem_bytes(J), stream>>>(p);
  else if (J == 19) gat_chain2_kernel<19><<<tiles, NT, smem_bytes(J), stream>>>(p);
  else gat_chain2_kernel<0><<<tiles, NT, smem_bytes(J), stream>>>(p);
  return check_launch("gat_chain2");
}

}  // namespace gator
