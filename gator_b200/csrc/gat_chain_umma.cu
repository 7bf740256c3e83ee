// Fused GAT blocks (lib/models/GAT.py:33-43 x depth) on tcgen05: ONE kernel runs all GATBlocks of the lifter for a
// tile of whole samples, with the residual stream in registers and every intermediate on the SM:
//
//   n  = LayerNorm1(x)
//   a  = proj(softmax_J(q k^T / 4 + hop_path_bias) v)        8 heads of 16, per sample        (modules.py:121-138)
//   g  = (A o I)(M o n W0) + (A o (1-I))(M o n W1) + b       modulated graph convolution      (modules.py:243-255)
//   x += linearback([1[hop<=1] L0(a+g) | 1[hop==2] L1(a+g)])                                  (modules.py:158-177)
//   x += fc2(GELU(fc1(LayerNorm2(x))))                                                        (modules.py:188-196)
//
// CTA = 128 token rows = S whole samples (S = 128 / J: 6 for J = 19, 7 for J = 17), 512 threads = 4 threads per
// row (32 of the 128 channels each; 16 of the 64 columns of a narrow unit).  Every dense product is a sequence of
// 36 "pieces" per block - 64 x 128 (N x K) or 128 x 64 bf16 hi/lo weight images (32 KB) streamed from L2 with
// cp.async into a 2-slot ring - accumulated in TMEM; the J x J mixes (attention, graph conv, hop masks) go through
// fp32 staging in shared memory between rows of the same sample.  x never leaves registers between blocks.
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
constexpr int PIECES = 36;
// piece indices inside a block's blob
enum { PC_QKV = 0, PC_PROJ = 8, PC_GCN = 10, PC_XF = 14, PC_MLP = 20 };

constexpr int A128_IMG = 128 * 128 * 2;        // 32 KB: 128 rows x K=128 image
constexpr int OFF_BUF1 = 0;                    // A operand, K = 128 layout (hi | lo)            65536
constexpr int OFF_BUF2 = 65536;                // A operand / fp32 staging                       65536
constexpr int OFF_W = 131072;                  // 2 weight slots                                 65536
constexpr int OFF_KB = 196608;                 // k of one head: [128][16] fp32                   8192
constexpr int OFF_VB = OFF_KB + 8192;          // v of one head                                   8192
constexpr int OFF_BIAS = OFF_VB + 8192;        // attention bias [8][J][J] fp32            <= 11552 (J = 19)
constexpr int smem_bytes(int J) { return OFF_BIAS + 8 * J * J * 4; }

struct GatChainParams {
  float* x;                       // (rows, 128) in/out
  int rows;                       // B * J
  int J, S, depth;
  const uint8_t* const* blobs;    // DEVICE array [depth] of piece blobs
  const float* const* prm;        // DEVICE array [depth * 14] of per-block fp32 parameter arrays (see PRM_*)
  const float* attn_bias;         // (8, J, J)
  const float* mask1;             // (J, J) 1[hop <= 1]
  const float* mask2;             // (J, J) 1[hop == 2]
  int split;
};
enum { PRM_LN1W = 0, PRM_LN1B, PRM_QKVB, PRM_PROJB, PRM_GCNM, PRM_ADIAG, PRM_AOFF, PRM_GCNB, PRM_XFB01, PRM_XFBB,
       PRM_LN2W, PRM_LN2B, PRM_FC1B, PRM_FC2B, PRM_COUNT };

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

__device__ __forceinline__ uint4 pk8(const float* v) {
  return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}
__device__ __forceinline__ uint4 pk8r(const float* v, const uint4& hi) {
  return make_uint4(pack_bf16(v[0] - bf16_lo_f(hi.x), v[1] - bf16_hi_f(hi.x)), pack_bf16(v[2] - bf16_lo_f(hi.y), v[3] - bf16_hi_f(hi.y)),
                    pack_bf16(v[4] - bf16_lo_f(hi.z), v[5] - bf16_hi_f(hi.z)), pack_bf16(v[6] - bf16_lo_f(hi.w), v[7] - bf16_hi_f(hi.w)));
}
// write NCH chunks (8 values each) of this row into an A image pair; KCH = chunks per row of the layout (16 or 8)
template <int NCH, int KCH>
__device__ __forceinline__ void write_a(uint8_t* buf, int row, int kc0, const float* v) {
  uint8_t* base = buf + (row >> 3) * (KCH * 128) + (row & 7) * 16;
  constexpr int LO = 128 * KCH * 16;           // size of the hi image
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const uint4 hi = pk8(v + 8 * c);
    *reinterpret_cast<uint4*>(base + (kc0 + c) * 128) = hi;
    *reinterpret_cast<uint4*>(base + LO + (kc0 + c) * 128) = pk8r(v + 8 * c, hi);
  }
}
__device__ __forceinline__ void ldg32(const float* src, float* dst) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(src) + i);
    dst[4 * i] = t.x; dst[4 * i + 1] = t.y; dst[4 * i + 2] = t.z; dst[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void ldg16(const float* src, float* dst) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(src) + i);
    dst[4 * i] = t.x; dst[4 * i + 1] = t.y; dst[4 * i + 2] = t.z; dst[4 * i + 3] = t.w;
  }
}

template <int JT>
__global__ void __launch_bounds__(NT, 1)
gat_chain_kernel(GatChainParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float2 xch[128][4];                   // LayerNorm partials [row][column quarter] (uses are separated by CTA barriers)
  __shared__ uint32_t hopbits[2][MAXJ];            // hop masks as bit sets per query joint
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int J = JT ? JT : p.J;
  constexpr int JU = JT ? JT : MAXJ;
  const int S = p.S;
  const int cq = warp >> 2;                        // column quarter: channels cq*32.. of a 128-wide row, cq*16.. of a unit
  const int row = (warp & 3) * 32 + lane;          // row in tile = TMEM lane
  const int row0 = blockIdx.x * S * J;             // first global row of the tile
  const int nrows = min(S * J, p.rows - row0);     // valid rows (whole samples)
  const bool valid = row < nrows;
  const int samp = valid ? row / J : 0;            // sample within the tile
  const int ji = valid ? row - samp * J : 0;       // joint index
  const int srow0 = samp * J;                      // first row of this row's sample
  uint8_t* buf1 = smem + OFF_BUF1;
  uint8_t* buf2 = smem + OFF_BUF2;
  float* kb = reinterpret_cast<float*>(smem + OFF_KB);
  float* vb = reinterpret_cast<float*>(smem + OFF_VB);
  float* sbias = reinterpret_cast<float*>(smem + OFF_BIAS);
  float* stage = reinterpret_cast<float*>(buf2);   // fp32 [128][128] (graph conv) or [128][64] at +32 KB (hop mix)

  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  if (tid == 32) {
    mbar_init(&bar, 1);
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

  int slot = 0;
  uint32_t phase = 0;
  auto prefetch = [&](const uint8_t* blob, int piece, int sl) {
    const uint8_t* src = blob + (size_t)piece * PIECE_BYTES;
    const uint32_t dst = smem_u32(smem + OFF_W + sl * PIECE_BYTES);
#pragma unroll
    for (int i = 0; i < PIECE_BYTES / 16 / NT; ++i) cp_async16(dst + (i * NT + tid) * 16, src + (i * NT + tid) * 16);
    cp_async_commit();
  };
  prefetch(p.blobs[0], 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t acc_u = tmem + lane_addr + cq * 16;       // this thread's 16 columns of the 64-wide unit accumulator
  const uint32_t acc_b = tmem + lane_addr + 64 + cq * 32;  // this thread's 32 columns of the 128-wide accumulator

  // One weight piece.  wide = false: D_unit[128 x 64] = A(K=128 layout, 8 k-steps) . W(64 x 128)^T
  //                    wide = true : D_big[128 x 128] (+)= A(4 k-steps from a_off, row-group stride a_sbo) . W(128 x 64)^T
  auto run_piece = [&](bool wide, uint32_t a_addr, uint32_t a_sbo, uint32_t a_lo, bool accumulate,
                       const uint8_t* next_blob, int next_piece) {
    cp_async_wait_all();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t w0 = smem_u32(smem + OFF_W + slot * PIECE_BYTES);
      const uint32_t d = wide ? tmem + 64 : tmem;
      const uint32_t idesc = wide ? idesc_bf16(128, 128) : idesc_bf16(128, 64);
      const uint32_t w_sbo = wide ? 1024u : 2048u;
      const int ksteps = wide ? 4 : 8;
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t ad = smem_desc(a_addr + ks * 256, 128, a_sbo), wd = smem_desc(w0 + ks * 256, 128, w_sbo);
        const uint32_t accf = (accumulate || ks > 0) ? 1u : 0u;
        if (p.split) {
          mma_bf16(d, smem_desc(a_addr + a_lo + ks * 256, 128, a_sbo), wd, idesc, accf);
          mma_bf16(d, ad, smem_desc(w0 + PIECE_IMG + ks * 256, 128, w_sbo), idesc, 1);
          mma_bf16(d, ad, wd, idesc, 1);
        } else {
          mma_bf16(d, ad, wd, idesc, accf);
        }
      }
      mma_commit(&bar);
    }
    slot ^= 1;
    if (next_blob) prefetch(next_blob, next_piece, slot);
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
  };
  // LayerNorm statistics over the 128 channels of a row held by 4 threads (32 each)
  auto stats128 = [&](const float* xr, float& mean, float& m2) {
    float m4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 32; ++i) m4[i & 3] += xr[i];
    const float m = ((m4[0] + m4[1]) + (m4[2] + m4[3])) * (1.0f / 32);
    float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 32; ++i) { const float d = xr[i] - m; q4[i & 3] = fmaf(d, d, q4[i & 3]); }
    xch[row][cq] = make_float2(m, (q4[0] + q4[1]) + (q4[2] + q4[3]));
    group_sync(1 + (warp & 3));
    const float2 a = xch[row][0], b = xch[row][1], c = xch[row][2], d = xch[row][3];
    mean = 0.25f * ((a.x + b.x) + (c.x + d.x));
    const float da = a.x - mean, db = b.x - mean, dc = c.x - mean, dd = d.x - mean;
    m2 = ((a.y + b.y) + (c.y + d.y)) + 32.0f * ((da * da + db * db) + (dc * dc + dd * dd));
  };

  const uint32_t b1 = smem_u32(buf1), b2 = smem_u32(buf2);
  float x[32], v[32], pw[32], pb[32];
  {
    const float4* src = reinterpret_cast<const float4*>(p.x + (size_t)(row0 + (valid ? row : 0)) * C + cq * 32);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 t = valid ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      x[4 * i] = t.x; x[4 * i + 1] = t.y; x[4 * i + 2] = t.z; x[4 * i + 3] = t.w;
    }
  }

#pragma unroll 1
  for (int blk = 0; blk < p.depth; ++blk) {
    const uint8_t* blob = p.blobs[blk];
    const uint8_t* nblob = blk + 1 < p.depth ? p.blobs[blk + 1] : nullptr;
    const float* const* prm = p.prm + blk * PRM_COUNT;
    // ---- LayerNorm1 -> buf1 (K = 128) ----
    {
      float mean, m2;
      stats128(x, mean, m2);
      const float rstd = rsqrtf(m2 * (1.0f / C) + 1e-5f);
      ldg32(prm[PRM_LN1W] + cq * 32, pw);
      ldg32(prm[PRM_LN1B] + cq * 32, pb);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = (x[i] - mean) * rstd * pw[i] + pb[i];
      write_a<4, 16>(buf1, row, cq * 4, v);
    }
    // ---- attention, one head per piece: unit = [q_h | k_h | v_h | 0] ----
#pragma unroll 1
    for (int h = 0; h < 8; ++h) {
      run_piece(false, b1, 2048, A128_IMG, false, blob, PC_QKV + h + 1);     // next: head h+1 or proj k-half 0
      float t16[16];
      tmem_ld16(acc_u, t16);
      tmem_ld_wait();
      if (cq < 3) {                                                          // + bias: q (0..127) | k (128..255) | v (256..383)
        float bq[16];
        ldg16(prm[PRM_QKVB] + cq * 128 + h * 16, bq);
#pragma unroll
        for (int i = 0; i < 16; ++i) t16[i] += bq[i];
      }
      if (cq == 1) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(kb + row * 16 + i) = make_float4(t16[i], t16[i + 1], t16[i + 2], t16[i + 3]);
      } else if (cq == 2) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(vb + row * 16 + i) = make_float4(t16[i], t16[i + 1], t16[i + 2], t16[i + 3]);
      }
      __syncthreads();
      if (cq == 0) {
        float s[JU];
        float mx = -INFINITY;
        const float* bias_row = sbias + (h * J + ji) * J;
#pragma unroll
        for (int j = 0; j < JU; ++j) {
          if (JT || j < J) {
            const float4* kr = reinterpret_cast<const float4*>(kb + (srow0 + j) * 16);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int d4 = 0; d4 < 4; ++d4) {
              const float4 kk = kr[d4];
              a0 = fmaf(t16[4 * d4], kk.x, a0); a1 = fmaf(t16[4 * d4 + 1], kk.y, a1);
              a2 = fmaf(t16[4 * d4 + 2], kk.z, a2); a3 = fmaf(t16[4 * d4 + 3], kk.w, a3);
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
        float o[16];
#pragma unroll
        for (int d = 0; d < 16; ++d) o[d] = 0.f;
#pragma unroll
        for (int j = 0; j < JU; ++j) {
          if (JT || j < J) {
            const float pj = s[j] * inv;
            const float4* vr = reinterpret_cast<const float4*>(vb + (srow0 + j) * 16);
#pragma unroll
            for (int d4 = 0; d4 < 4; ++d4) {
              const float4 vv = vr[d4];
              o[4 * d4] = fmaf(pj, vv.x, o[4 * d4]); o[4 * d4 + 1] = fmaf(pj, vv.y, o[4 * d4 + 1]);
              o[4 * d4 + 2] = fmaf(pj, vv.z, o[4 * d4 + 2]); o[4 * d4 + 3] = fmaf(pj, vv.w, o[4 * d4 + 3]);
            }
          }
        }
        write_a<2, 16>(buf2, row, h * 2, o);
      }
    }
    // ---- proj: a = o Wp^T (2 K-halves) ----
    run_piece(true, b2, 2048, A128_IMG, false, blob, PC_PROJ + 1);
    run_piece(true, b2 + 1024, 2048, A128_IMG, true, blob, PC_GCN);
    float s_[32];                                     // s = a + g, built up in registers
    tmem_ld32(acc_b, s_);
    tmem_ld_wait();
    ldg32(prm[PRM_PROJB] + cq * 32, pw);
#pragma unroll
    for (int i = 0; i < 32; ++i) s_[i] += pw[i];
    // ---- modulated graph convolution: h0 = n W0, h1 = n W1 ----
    run_piece(true, b1, 2048, A128_IMG, false, blob, PC_GCN + 1);
    run_piece(true, b1 + 1024, 2048, A128_IMG, true, blob, PC_GCN + 2);
    tmem_ld32(acc_b, v);
    tmem_ld_wait();
    ldg32(prm[PRM_GCNM] + ji * C + cq * 32, pw);      // M[ji, channels]
    {
      const float ad = __ldg(prm[PRM_ADIAG] + ji);
      ldg32(prm[PRM_GCNB] + cq * 32, pb);
#pragma unroll
      for (int i = 0; i < 32; ++i) s_[i] += ad * (pw[i] * v[i]) + pb[i];
    }
    run_piece(true, b1, 2048, A128_IMG, false, blob, PC_GCN + 3);
    run_piece(true, b1 + 1024, 2048, A128_IMG, true, blob, PC_XF);
    tmem_ld32(acc_b, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; i += 4)                   // stage M o h1 for the other joints of the sample (buf2: o is dead)
      *reinterpret_cast<float4*>(stage + row * C + cq * 32 + i) = make_float4(pw[i] * v[i], pw[i + 1] * v[i + 1], pw[i + 2] * v[i + 2], pw[i + 3] * v[i + 3]);
    __syncthreads();
    {
      const float* aoff = prm[PRM_AOFF] + ji * J;
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
      for (int j = 0; j < JU; ++j) {
        if (JT || j < J) {
          const float aij = __ldg(aoff + j);
          const float4* sr = reinterpret_cast<const float4*>(stage + (srow0 + j) * C + cq * 32);
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            const float4 t = sr[i4];
            v[4 * i4] = fmaf(aij, t.x, v[4 * i4]); v[4 * i4 + 1] = fmaf(aij, t.y, v[4 * i4 + 1]);
            v[4 * i4 + 2] = fmaf(aij, t.z, v[4 * i4 + 2]); v[4 * i4 + 3] = fmaf(aij, t.w, v[4 * i4 + 3]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) s_[i] += v[i];
    }
    write_a<4, 16>(buf1, row, cq * 4, s_);            // n is dead: buf1 <- s
    __syncthreads();                                   // every reader of `stage` is done before buf2 is reused
    // ---- X_Feat: 3 units of L01 (64 columns each) -> hop mix -> linearback K-slices ----
#pragma unroll 1
    for (int u = 0; u < 3; ++u) {
      run_piece(false, b1, 2048, A128_IMG, false, blob, PC_XF + 2 * u + 1);
      float y[16];
      tmem_ld16(acc_u, y);
      tmem_ld_wait();
      {
        float bq[16];
        ldg16(prm[PRM_XFB01] + u * 64 + cq * 16, bq);
        float* st = stage + 8192 + row * 64 + cq * 16;        // fp32 [128][64] in the upper half of buf2
#pragma unroll
        for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(st + i) = make_float4(y[i] + bq[i], y[i + 1] + bq[i + 1], y[i + 2] + bq[i + 2], y[i + 3] + bq[i + 3]);
      }
      __syncthreads();
      {
        const uint32_t bits = hopbits[u == 2 ? 1 : 0][ji];
#pragma unroll
        for (int i = 0; i < 16; ++i) y[i] = 0.f;
#pragma unroll
        for (int j = 0; j < JU; ++j) {
          if ((JT || j < J) && ((bits >> j) & 1u)) {
            const float4* sr = reinterpret_cast<const float4*>(stage + 8192 + (srow0 + j) * 64 + cq * 16);
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 t = sr[i4];
              y[4 * i4] += t.x; y[4 * i4 + 1] += t.y; y[4 * i4 + 2] += t.z; y[4 * i4 + 3] += t.w;
            }
          }
        }
        write_a<2, 8>(buf2, row, cq * 2, y);                  // K = 64 operand in the lower half of buf2
      }
      run_piece(true, b2, 1024, 128 * 64 * 2, u > 0, blob, u < 2 ? PC_XF + 2 * u + 2 : PC_MLP);
    }
    tmem_ld32(acc_b, v);
    tmem_ld_wait();
    ldg32(prm[PRM_XFBB] + cq * 32, pw);
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] += v[i] + pw[i];
    // ---- MLP ----
    {
      float mean, m2;
      stats128(x, mean, m2);
      const float rstd = rsqrtf(m2 * (1.0f / C) + 1e-5f);
      ldg32(prm[PRM_LN2W] + cq * 32, pw);
      ldg32(prm[PRM_LN2B] + cq * 32, pb);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = (x[i] - mean) * rstd * pw[i] + pb[i];
      write_a<4, 16>(buf1, row, cq * 4, v);
    }
#pragma unroll 1
    for (int u = 0; u < 8; ++u) {
      run_piece(false, b1, 2048, A128_IMG, false, blob, PC_MLP + 2 * u + 1);
      float y[16], bq[16];
      tmem_ld16(acc_u, y);
      tmem_ld_wait();
      ldg16(prm[PRM_FC1B] + u * 64 + cq * 16, bq);
#pragma unroll
      for (int i = 0; i < 16; ++i) y[i] = gelu_erf_fast(y[i] + bq[i]);
      write_a<2, 8>(buf2, row, cq * 2, y);
      const bool last = u == 7;
      run_piece(true, b2, 1024, 128 * 64 * 2, u > 0, last ? nblob : blob, last ? 0 : PC_MLP + 2 * u + 2);
    }
    tmem_ld32(acc_b, v);
    tmem_ld_wait();
    ldg32(prm[PRM_FC2B] + cq * 32, pw);
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] += v[i] + pw[i];
  }
  if (valid) {
    float4* dst = reinterpret_cast<float4*>(p.x + (size_t)(row0 + row) * C + cq * 32);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace

bool gat_chain_supported(int J) { return J >= 2 && J <= 21; }   // attention-bias table [8][J][J] must fit beside the operands

int launch_gat_chain(float* x, int rows, int J, int depth, const void* const* blobs_dev, const float* const* prm_dev,
                     const float* attn_bias, const float* mask1, const float* mask2, bool split, cudaStream_t stream) {
  static DeviceOnce attr_once;
  GATOR_TRY(attr_once.run("gat_chain", [&](int) -> cudaError_t {
    GATOR_CUDA_OK(cudaFuncSetAttribute(gat_chain_kernel<17>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(17)));
    GATOR_CUDA_OK(cudaFuncSetAttribute(gat_chain_kernel<19>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(19)));
    return cudaFuncSetAttribute(gat_chain_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(21));
  }));
  GATOR_REQUIRE(gat_chain_supported(J), "gat_chain: num_joint=%d does not fit the fused kernel", J);
  GatChainParams p;
  p.x = x; p.rows = rows; p.J = J; p.S = 128 / J; p.depth = depth;
  p.blobs = reinterpret_cast<const uint8_t* const*>(blobs_dev);
  p.prm = prm_dev; p.attn_bias = attn_bias; p.mask1 = mask1; p.mask2 = mask2; p.split = split ? 1 : 0;
  const int rows_per_tile = p.S * J;
  const int tiles = (rows + rows_per_tile - 1) / rows_per_tile;
  if (J == 17) gat_chain_kernel<17><<<tiles, NT, smem_bytes(J), stream>>>(p);
  else if (J == 19) gat_chain_kernel<19><<<tiles, NT, smem_bytes(J), stream>>>(p);
  else gat_chain_kernel<0><<<tiles, NT, smem_bytes(J), stream>>>(p);
  return check_launch("gat_chain");
}

}  // namespace gator
