// SMPL skinning on the tensor cores (smplpytorch/smplpytorch/pytorch/smpl_layer.py:134-145), GATOR_PREC_BF16X3 only.
//
// The reference blends the joint transforms with a dense product th_T = th_results2 @ th_weights^T (:134) and applies
// them per vertex (:136-145).  Here the blend is that same product as a tcgen05 GEMM
//     T[v, (s, e)] = sum_j W[v, j] * G'[s, j, e]          v: vertex, s: sample, e: entry of the 3x4 transform, j: joint
// with M = 128 vertices per tile, N = 240 = 20 samples x 12 entries, K = 32 (24 joints, zero padded), 3-term bf16 split,
// and the apply step out[s, v, :] = (T[v, s] [p; 1] + offset_s) * scale runs in the epilogue straight from TMEM:
// 12 accumulator values, 3 loads and 3 stores per (vertex, sample) instead of 48 shared-memory gathers + 48 FMAs
// (the CUDA-core kernel is bound by those gathers: 48 conflict-free wavefronts per warp and sample).
//
// Same skeleton as csrc/umma_gemm_wide.cu: persistent CTAs, warp 0 = TMA producer (one 16 KB weight tile + one 32 KB
// transform tile per stage - K fits one block - and, in a ring of its own, the tile's v_posed block: one 1536-byte bulk
// copy per sample row, 20 x 128 vertices), warp 1 = MMA issue into one of two 256-column TMEM accumulators,
// warps 2-17 = epilogue (four warps per TMEM lane quarter, 5 samples each): lane = vertex, its rest position read from
// the v_posed block with a stride of 3 words (conflict free), the skinned vertices staged in warp-private shared memory
// and stored as 256-byte coalesced rows of 8-byte accesses (a sample row is 8 mod 16 bytes).
// Round 1 loaded v_posed with per-lane global loads and transposed it through shared memory in the epilogue warps: 103
// instructions per (vertex, sample), issue slots 67 % busy, 286 us per 8192 samples = 70 % of the HBM roofline of its
// own traffic (profiles/r02_ncu_smpl_kernels.txt).
#include "common.cuh"
#include "umma.cuh"

namespace gator {
namespace {

using namespace umma;

constexpr int NJ = GATOR_SMPL_JOINTS;   // 24
constexpr int NV = GATOR_V_FULL;        // 6890
constexpr int NV3 = NV * 3;
constexpr int TM = 128;                 // vertices per tile
constexpr int TS = 20;                  // samples per tile
constexpr int TN = TS * 12;             // 240 accumulator columns
constexpr int A_HALF = TM * 32 * 2;     // 8 KB: hi or lo of a weight tile (128 x 32 bf16)
constexpr int B_HALF = 256 * 32 * 2;    // 16 KB: hi or lo of a transform tile (256 rows, 240 used)
constexpr int A_TILE = 2 * A_HALF, B_TILE = 2 * B_HALF;
constexpr int STAGE = A_TILE + B_TILE;  // 48 KB
constexpr int STAGES = 2;
constexpr int P_ROW = TM * 3 * 4;       // 1536 B: the v_posed floats of a tile's 128 vertices, one sample
constexpr int P_STAGE = TS * P_ROW;     // 30 KB
constexpr int P_STAGES = 2;
constexpr int GROUPS = 4;               // epilogue warps per TMEM lane quarter (5 measured no faster)
constexpr int HS = TS / GROUPS;         // samples per epilogue warp and tile (5)
constexpr int EPI_WARPS = 4 * GROUPS;
constexpr int NTHREADS = (2 + EPI_WARPS) * 32;
constexpr int OFF_P = STAGES * STAGE;
constexpr int OFF_ST = OFF_P + P_STAGES * P_STAGE;
constexpr int SMEM = OFF_ST + EPI_WARPS * HS * 96 * 4;            // + [warp][sample][96] output staging (30 KB)

struct SkinParams {
  const uint8_t* Wimg;     // [m_tiles][hi|lo][16][4][8][8]
  const uint8_t* Timg;     // [n_tiles][hi|lo][32][4][8][8]
  const float* vposed;     // (S, ld)
  const float* offset;     // (S, 3)
  float* verts;            // (S, 6890, 3)
  int S, ld, m_tiles, n_tiles;
  float scale;
};

// G' (S, 24, 12) fp32 -> transform tile image: row n = s_local * 12 + e, k = joint.  One thread = one 16-byte chunk.
__global__ void __launch_bounds__(256)
skin_t_image_kernel(const float* __restrict__ amat, int S, long long chunks, uint8_t* __restrict__ img) {
  const long long c = (long long)blockIdx.x * 256 + threadIdx.x;
  if (c >= chunks) return;
  const int r = c & 7, kc = (c >> 3) & 3, rg = (c >> 5) & 31;
  const int nt = (int)(c >> 10);
  const int n = rg * 8 + r;                      // row in the tile
  const int sl = n / 12, e = n - sl * 12;
  const int s = nt * TS + sl;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int j = kc * 8 + i;
    v[i] = (n < TN && s < S && j < NJ) ? amat[((size_t)s * NJ + j) * 12 + e] : 0.f;
  }
  const float4 a = make_float4(v[0], v[1], v[2], v[3]), b = make_float4(v[4], v[5], v[6], v[7]);
  const uint4 hi = cvt8(a, b);
  uint4* dst = reinterpret_cast<uint4*>(img + (size_t)nt * B_TILE) + (c & 1023);
  dst[0] = hi;
  dst[B_HALF / 16] = cvt8_residual(a, b, hi);
}

__global__ void __launch_bounds__(NTHREADS, 1)
smpl_skin_umma_kernel(SkinParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[STAGES], empty[STAGES], acc_full[2], acc_empty[2], p_full[P_STAGES], p_empty[P_STAGES];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < P_STAGES; ++s) { mbar_init(&p_full[s], 1); mbar_init(&p_empty[s], EPI_WARPS); }
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], EPI_WARPS * 32); mbar_init(&acc_empty[1], EPI_WARPS * 32);
    mbar_init_fence();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int total = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int nt = t / p.m_tiles, mt = t - nt * p.m_tiles;
        const int s = it % STAGES;
        if (it >= STAGES) mbar_wait(&empty[s], ((it / STAGES) - 1) & 1);
        mbar_arrive_expect_tx(&full[s], STAGE);
        bulk_copy_g2s(smem + s * STAGE, p.Wimg + (size_t)mt * A_TILE, A_TILE, &full[s]);
        bulk_copy_g2s(smem + s * STAGE + A_TILE, p.Timg + (size_t)nt * B_TILE, B_TILE, &full[s]);
        // the tile's v_posed block: rows of 128 vertices (the last vertex tile stops at the padded end of the row)
        const int ps = it % P_STAGES;
        if (it >= P_STAGES) mbar_wait(&p_empty[ps], ((it / P_STAGES) - 1) & 1);
        const int ns = min(TS, p.S - nt * TS);
        const int row_bytes = min(P_ROW, p.ld * 4 - mt * P_ROW);
        mbar_arrive_expect_tx(&p_full[ps], (uint32_t)(ns * row_bytes));
        const float* src = p.vposed + (size_t)nt * TS * p.ld + (size_t)mt * (P_ROW / 4);
        uint8_t* dst = smem + OFF_P + ps * P_STAGE;
        for (int i = 0; i < ns; ++i) bulk_copy_g2s(dst + i * P_ROW, src + (size_t)i * p.ld, (uint32_t)row_bytes, &p_full[ps]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16(TM, TN);
      int it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int buf = it & 1, s = it % STAGES;
        if (it >= 2) mbar_wait(&acc_empty[buf], ((it >> 1) - 1) & 1);
        mbar_wait(&full[s], (it / STAGES) & 1);
        tc_fence_after();
        const uint32_t d = tmem + buf * 256;
        const uint32_t a0 = smem_u32(smem + s * STAGE), b0 = a0 + A_TILE;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t ad = smem_desc(a0 + ks * 256, 128, 512), adl = smem_desc(a0 + A_HALF + ks * 256, 128, 512);
          const uint64_t bd = smem_desc(b0 + ks * 256, 128, 512), bdl = smem_desc(b0 + B_HALF + ks * 256, 128, 512);
          mma_bf16(d, adl, bd, idesc, ks);
          mma_bf16(d, ad, bdl, idesc, 1);
          mma_bf16(d, ad, bd, idesc, 1);
        }
        mma_commit(&empty[s]);
        mma_commit(&acc_full[buf]);
      }
    }
  } else {
    const int ew = warp - 2;
    const int q = warp & 3;                        // TMEM lane quarter = which 32 vertices of the tile
    const int grp = ew >> 2;                       // which HS of the tile's samples
    float* st = reinterpret_cast<float*>(smem + OFF_ST) + ew * HS * 96;
    // per-sample offsets of a tile (one element per lane) are requested one tile ahead
    float ofs = 0.f;
    auto request = [&](int t) {
      const int nt = t / p.m_tiles;
      const int s_lo = nt * TS + grp * HS;
      const int ns = max(0, min(HS, p.S - s_lo));
      ofs = lane < ns * 3 ? p.offset[(size_t)s_lo * 3 + lane] : 0.f;
    };
    if ((int)blockIdx.x < total) request(blockIdx.x);
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int nt = t / p.m_tiles, mt = t - nt * p.m_tiles;
      const int buf = it & 1, ps = it % P_STAGES;
      const int vbase = mt * TM + q * 32;
      const int nf = max(0, min(32, NV - vbase)) * 3;            // floats of this warp's vertex segment
      const int s_lo = nt * TS + grp * HS;
      const int ns = max(0, min(HS, p.S - s_lo));
      // rest positions: lane = vertex, stride of 3 words (rows past the batch / vertices past the mesh hold stale data
      // that is computed on but never stored)
      mbar_wait(&p_full[ps], (it / P_STAGES) & 1);
      const float* pt = reinterpret_cast<const float*>(smem + OFF_P + ps * P_STAGE) + (grp * HS) * (P_ROW / 4) + q * 96 + lane * 3;
      float pv[HS][3];
#pragma unroll
      for (int i = 0; i < HS; ++i) { pv[i][0] = pt[i * (P_ROW / 4)]; pv[i][1] = pt[i * (P_ROW / 4) + 1]; pv[i][2] = pt[i * (P_ROW / 4) + 2]; }
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_empty[ps]);
      const float ofs_cur = ofs;
      if (t + (int)gridDim.x < total) request(t + gridDim.x);
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      if (nf > 0 && ns > 0) {
        const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + buf * 256 + grp * HS * 12;
        // one sample at a time (16 registers of accumulator columns)
#pragma unroll
        for (int i = 0; i < HS; ++i) {
          if (i < ns) {
            float tv[16];
            tmem_ld16(tbase + i * 12, tv);                       // 16 columns, 12 used
            const float o0 = __shfl_sync(0xffffffffu, ofs_cur, i * 3), o1 = __shfl_sync(0xffffffffu, ofs_cur, i * 3 + 1),
                        o2 = __shfl_sync(0xffffffffu, ofs_cur, i * 3 + 2);
            const float px = pv[i][0], py = pv[i][1], pz = pv[i][2];
            tmem_ld_wait();
            st[i * 96 + lane * 3] = (tv[0] * px + tv[1] * py + tv[2] * pz + tv[3] + o0) * p.scale;
            st[i * 96 + lane * 3 + 1] = (tv[4] * px + tv[5] * py + tv[6] * pz + tv[7] + o1) * p.scale;
            st[i * 96 + lane * 3 + 2] = (tv[8] * px + tv[9] * py + tv[10] * pz + tv[11] + o2) * p.scale;
          }
        }
        __syncwarp();
        // 8-byte accesses: a sample row is 82 680 B = 8 mod 16, a segment starts a multiple of 384 B into it
        const int nf2 = nf >> 1;
        float2* dst0 = reinterpret_cast<float2*>(p.verts + (size_t)s_lo * NV3 + (size_t)vbase * 3);
        const float2* st2 = reinterpret_cast<const float2*>(st);
#pragma unroll
        for (int i = 0; i < HS; ++i) {
          if (i < ns) {
            float2* dst = dst0 + (size_t)i * (NV3 / 2);
            if (lane < nf2) dst[lane] = st2[i * 48 + lane];
            if (lane + 32 < nf2) dst[lane + 32] = st2[i * 48 + lane + 32];
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace

size_t skin_t_image_bytes(int S) { return (size_t)((S + TS - 1) / TS) * B_TILE; }

// verts = (blend(W, G') [v_posed; 1] + offset) * scale for S samples.  Wimg: skin weights as the 128-row tile image
// (gator_b200/packing.py: pack_umma_wide_a); timg: workspace of skin_t_image_bytes(S).
int launch_smpl_skin_umma(const float* vposed, int ld, const float* amat, const float* offset, const void* Wimg, void* timg,
                          float* verts, int S, float scale, cudaStream_t stream) {
  if (S <= 0) return GATOR_OK;
  GATOR_REQUIRE(ld % 4 == 0 && ld >= NV3 && (reinterpret_cast<uintptr_t>(vposed) & 15u) == 0,
                "smpl_skin_umma: v_posed rows must be 16-byte aligned (bulk copies) and hold 20670 floats");
  SkinParams p;
  p.Wimg = static_cast<const uint8_t*>(Wimg);
  p.Timg = static_cast<const uint8_t*>(timg);
  p.vposed = vposed; p.offset = offset; p.verts = verts;
  p.S = S; p.ld = ld; p.scale = scale;
  p.m_tiles = (NV + TM - 1) / TM;
  p.n_tiles = (S + TS - 1) / TS;
  const long long chunks = (long long)p.n_tiles * 1024;
  skin_t_image_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, stream>>>(amat, S, chunks, static_cast<uint8_t*>(timg));
  GATOR_TRY(check_launch("skin_t_image"));
  static DeviceOnce attr_once;
  static int sm_count[64];
  int dev = 0;
  cudaGetDevice(&dev);
  GATOR_TRY(attr_once.run("smpl_skin_umma", [&](int d) -> cudaError_t {
    GATOR_CUDA_OK(cudaFuncSetAttribute(smpl_skin_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    return cudaDeviceGetAttribute(&sm_count[d & 63], cudaDevAttrMultiProcessorCount, d);
  }));
  const int sms = sm_count[dev & 63] > 0 ? sm_count[dev & 63] : 148;
  const int total = p.m_tiles * p.n_tiles;
  smpl_skin_umma_kernel<<<total < sms ? total : sms, NTHREADS, SMEM, stream>>>(p);
  return check_launch("smpl_skin_umma");
}

}  // namespace gator
