/*
 * gator_b200 - C ABI of the B200-native (sm_100a) GATOR pose->mesh forward path.
 *
 * The reference (kasvii/GATOR) is pure Python/PyTorch and has no FFI of its own; every entry point
 * below replaces the ATen dispatch sequence of one reference `forward` (file:line relative to the
 * reference root).  The host side (gator_b200/*.py) mirrors the reference's nn.Module interface and
 * binds these symbols with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the name ends in _host;
 *   - every buffer (inputs, packed weights, workspace, outputs) is allocated and owned by the caller
 *     (PyTorch); the library never allocates or frees device memory and keeps no pointer after return;
 *   - `stream` is a cudaStream_t passed as void*; no call synchronises the host, so every forward is
 *     CUDA-graph capturable;
 *   - return value: 0 on success, a negative gator_status otherwise; gator_last_error() gives a
 *     thread-local message.  The library never throws and never exits.
 *   - weights are passed as a table of pointers indexed by the slot enums below; the slot NAMES are
 *     exported (gator_*_slot_name) so the Python packer binds by name, not by number.
 */
#ifndef GATOR_B200_H_
#define GATOR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GATOR_ABI_VERSION 2

typedef enum {
  GATOR_OK = 0,
  GATOR_ERR_BAD_ARG = -1,      /* null pointer, bad shape, misaligned buffer */
  GATOR_ERR_WORKSPACE = -2,    /* workspace too small */
  GATOR_ERR_LAUNCH = -3,       /* cudaGetLastError() after a launch */
  GATOR_ERR_ARCH = -4,         /* device is not sm_100 */
  GATOR_ERR_UNSUPPORTED = -5
} gator_status;

/* precision of the matrix products (accumulation is always fp32) */
typedef enum {
  GATOR_PREC_FP32 = 0,         /* FFMA everywhere: the <=1e-4 m parity path */
  GATOR_PREC_BF16 = 1,         /* tcgen05, single bf16 product per term                    */
  GATOR_PREC_BF16X3 = 2        /* tcgen05, 3-term bf16 split in the GEMMs (a_hi w_hi + a_lo w_hi + a_hi w_lo,
                                  ~2^-17 relative), bf16 operands in the attention cores: the default fast path */
} gator_precision;

int gator_abi_version(void);
const char* gator_last_error(void);
/* sizeof() of the ABI structs as compiled, so the ctypes mirror can be verified at load time:
 * which = 0 gat, 1 mdr, 2 smpl, 3 csr, 4 gemm, 5 eval, 6 pose2d, 7 smpl_cam, 8 upsample2, 9 lbs, 10 mano_post */
size_t gator_abi_sizeof(int which);
/* number of kernels this library has launched (process-wide); reset != 0 zeroes it after reading.
 * bench.py reports it as `gpu_launches`. */
long long gator_launch_count(int reset);

/* ------------------------------------------------------------------------------------------------
 * GAT lifter - replaces GAT.forward (lib/models/GAT.py:133-152) incl. GATBlock.forward (:33-43),
 * Attention/MGCN/X_Feat/MLP (lib/models/backbones/modules.py:121-138,243-255,158-177,188-196),
 * GraphLinear+GroupNorm embedding (GAT.py:69-72,135-144).  embed_dim=128, heads=8 (all call sites).
 * ---------------------------------------------------------------------------------------------- */
typedef enum {
  /* global */
  GAT_EMB_W1 = 0,     /* (64,2)    GLinear.0.W                                   */
  GAT_EMB_B1,         /* (64)      GLinear.0.b                                   */
  GAT_GN_W,           /* (64)      GLinear.1.weight  (GroupNorm(4,64))            */
  GAT_GN_B,           /* (64)      GLinear.1.bias                                */
  GAT_EMB_W2T,        /* (64,128)  GLinear.3.W transposed                        */
  GAT_EMB_B2,         /* (128)     GLinear.3.b                                   */
  GAT_POS_CONST,      /* (J,128)   pos_id_embed(1..J) + pos_num_embed(row degree) */
  GAT_ATTN_BIAS,      /* (8,J,J)   HopPathEncoding.forward() (modules.py:98-107)  */
  GAT_HOP_MASK1,      /* (J,J)     1[hop<=1]  (modules.py:165-166)                */
  GAT_HOP_MASK2,      /* (J,J)     1[hop==2]  (modules.py:167-168)                */
  GAT_NORM_W,         /* (128)     norm.weight                                   */
  GAT_NORM_B,         /* (128)                                                    */
  GAT_LIFT_W,         /* (3J,128J) lifter.weight                                 */
  GAT_LIFT_B,         /* (3J)                                                     */
  GAT_CHAIN_BLOBS,    /* fused-blocks kernel (csrc/gat_chain2_umma.cu), may be NULL: DEVICE array [depth] of pointers to
                         34 x 32 KB bf16 hi|lo tcgen05 weight pieces per block, in the order the kernel consumes them:
                         q rows 0-63, 64-127, k heads 0-3, v heads 0-3, k heads 4-7, v heads 4-7 (64x128 each);
                         proj K-halves (128x64) x2; gcn W0 K-halves x2, W1 K-halves x2; [L0;L1] rows 64u.. (64x128) x3;
                         linearback[:, 64u..] (128x64) x3; fc1 rows 64u.. u=0..3; fc2[:, 64u..] u=0..3; fc1 u=4..7; fc2 u=4..7 */
  GAT_CHAIN_PRM,      /* DEVICE array [depth*14] of fp32 arrays: LN1_W, LN1_B, QKV_B, PROJ_B, GCN_M, GCN_ADIAG, GCN_AOFF,
                         GCN_BIAS, XF_B01 (zero-padded to 192), XF_BB, LN2_W, LN2_B, FC1_B, FC2_B; may be NULL */
  GAT_NUM_GLOBAL
} gator_gat_global_slot;

typedef enum {
  GATB_LN1_W = 0, GATB_LN1_B,       /* (128)                                          */
  GATB_QKV_W, GATB_QKV_B,           /* (384,128),(384)                                */
  GATB_PROJ_W, GATB_PROJ_B,         /* (128,128),(128)                                */
  GATB_GCN_W01,                     /* (256,128) rows 0..127 = gcn.W[0]^T, 128..255 = gcn.W[1]^T */
  GATB_GCN_M,                       /* (J,128)                                        */
  GATB_GCN_ADIAG,                   /* (J)   diag of ((adj+adj2)+(adj+adj2)^T)/2       */
  GATB_GCN_AOFF,                    /* (J,J) same matrix with the diagonal zeroed     */
  GATB_GCN_BIAS,                    /* (128)                                          */
  GATB_XF_W01, GATB_XF_B01,         /* (144,128),(144): x_feat.linears.0 ; linears.1   */
  GATB_XF_WB, GATB_XF_BB,           /* (128,144),(128): x_feat.linearback             */
  GATB_LN2_W, GATB_LN2_B,           /* (128)                                          */
  GATB_FC1_W, GATB_FC1_B,           /* (512,128),(512)                                */
  GATB_FC2_W, GATB_FC2_B,           /* (128,512),(128)                                */
  GATB_NUM
} gator_gat_block_slot;

typedef struct {
  int32_t num_joint;           /* J: 17 or 19 (any 2..32)                              */
  int32_t depth;               /* number of GATBlocks (6 at every call site)           */
  int32_t batch;               /* B                                                    */
  int32_t chunk;               /* samples per pass through the workspace (0 = default) */
  int32_t precision;           /* gator_precision                                      */
  int32_t reserved;
  const void* const* weights;  /* HOST array of GAT_NUM_GLOBAL + depth*GATB_NUM device pointers */
  const void* const* weights_bf16; /* same indexing: tcgen05-packed bf16 copy of each *_W matrix slot (see
                                  gator_umma_weight_layout), NULL entries / NULL table = fp32 kernel */
  const void* const* weights_bf16_lo; /* packed bf16 residuals W - bf16(W) for GATOR_PREC_BF16X3 (may be NULL) */
  const float* pose2d;         /* (B,J,2)                                              */
  float* pose3d;               /* (B,3J)   x_out, millimetres                          */
  float* feat;                 /* (B,J,128) GELU(LN(x)) - second return of GAT.forward */
  void* workspace;
  size_t workspace_bytes;
} gator_gat_args;

const char* gator_gat_slot_name(int slot);        /* slot < GAT_NUM_GLOBAL: global; else block slot */
size_t gator_gat_workspace_bytes(int32_t batch, int32_t num_joint, int32_t chunk);
int gator_gat_forward(const gator_gat_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MDR decoder - replaces MDR.forward (lib/models/MDR.py:124-170) incl. CrossAttentionBlock (:64-69),
 * CrossAttention (:34-46), LayerNorm / MultiHeadedAttention
 * (lib/models/vanilla_transformer_encoder.py:24-46,82-94), the MDR head (:156-166) and
 * upsample_conv + template (:167-168).  The concat of GATOR.forward (GATOR.py:19) is folded in:
 * the kernel reads pose2d, pose3d (mm) and feat separately.
 * ---------------------------------------------------------------------------------------------- */
typedef enum {
  MDR_JF_WFEAT = 0,   /* (64,128)   get_joint_feature.weight[:, 5:]                          */
  MDR_JF_WPOSE,       /* (64,5)     get_joint_feature.weight[:, :5]                          */
  MDR_JF_BIASROWS,    /* (J,64)     get_joint_feature.bias + pos_j_id_embed(1..J)            */
  MDR_VF_CONST,       /* (431,64)   W_v[:, :3] @ init_vertices^T + b_v + pos_v_id_embed(1..431) */
  MDR_VF_W3,          /* (64,3)     get_verts_feature.weight[:, 3:6]                         */
  MDR_VJ,             /* (431) int32 nearest-joint table (graph_utils.py:71-89)              */
  MDR_HEAD_W,         /* (28,64)    rows 0..22 motion_linear, 23..25 bias_linear, 26 scale_linear, 27 zero */
  MDR_HEAD_B,         /* (28)                                                                 */
  MDR_BNORM_SCALE,    /* alpha=0: (431) w/sqrt(rv+eps); alpha=1: (3) LayerNorm(3).weight      */
  MDR_BNORM_SHIFT,    /* alpha=0: (431) b - rm*scale;   alpha=1: (3) LayerNorm(3).bias        */
  MDR_BCONV_W,        /* (20,431,3) bias_conv1d.weight                                       */
  MDR_BCONV_B,        /* (20)                                                                 */
  MDR_UP_W,           /* (6890,1296) upsample_conv.weight flattened (431*3=1293), K zero-padded */
  MDR_UP_BIAST,       /* (6890,3)   upsample_conv.bias[:,None] + init_vertices_6890            */
  MDR_CHAIN_FINAL,    /* bf16 blob for the final pass of the fused layer kernel: 2 units x [hi 8 KB | lo 8 KB]: the last
                         layer's selfatt.linears.3 and HEAD_W zero-padded to 64 rows; may be NULL          */
  MDR_UP_W_WIDE,      /* UP_W as the tile-major split image of the wide-N kernel (gator_umma_wide_layout);
                         used by GATOR_PREC_BF16X3; may be NULL (then the generic tcgen05 GEMM runs)       */
  MDR_NUM_GLOBAL
} gator_mdr_global_slot;

typedef enum {
  MDRL_N1_W = 0, MDRL_N1_B,         /* (64) encoder.norm1                                 */
  MDRL_WQ,                          /* (64,64)  encoder.attn.wq.weight                    */
  MDRL_WKV,                         /* (128,64) wk ; wv                                   */
  MDRL_PROJ_W, MDRL_PROJ_B,         /* (64,64),(64)                                       */
  MDRL_N2_W, MDRL_N2_B,             /* (64) encoder.norm2                                 */
  MDRL_FC1_W, MDRL_FC1_B,           /* (256,64),(256)                                     */
  MDRL_FC2_W, MDRL_FC2_B,           /* (64,256),(64)                                      */
  MDRL_CLN_A, MDRL_CLN_B,           /* (64) norm.a_2, norm.b_2 (unbiased-std LayerNorm)   */
  MDRL_SQKV_W, MDRL_SQKV_B,         /* (192,64),(192) selfatt.linears.0 ; 1 ; 2           */
  MDRL_SO_W, MDRL_SO_B,             /* (64,64),(64)   selfatt.linears.3                   */
  MDRL_CHAIN,                       /* bf16 blob for the fused layer kernel: 14 units x [hi 8 KB | lo 8 KB] of 64x64
                                       tcgen05 images: previous layer's linears.3 (zeros for layer 0), wq, proj,
                                       fc1 row quarters x4, fc2 column quarters x4, selfatt linears.0/1/2; may be NULL */
  MDRL_NUM
} gator_mdr_layer_slot;

#define GATOR_MDR_LAYERS 3
#define GATOR_V_FULL 6890
#define GATOR_V_COARSE 431
#define GATOR_UP_K 1296

typedef struct {
  int32_t num_joint;           /* J                                                     */
  int32_t batch;               /* B                                                     */
  int32_t chunk;               /* samples per pass (0 = default)                        */
  int32_t alpha;               /* cfg.MODEL.alpha (MDR.py:115,162)                      */
  int32_t precision;           /* gator_precision                                       */
  int32_t reserved;
  int32_t pose3d_metres;       /* 0: pose3d is in millimetres and divided by 1000 inside (GATOR.py:19);
                                  1: pose3d is already in metres (MDR.forward's pose_combine[:, :, 2:5], MDR.py:127)  */
  int32_t reserved2;
  const void* const* weights;  /* HOST array of MDR_NUM_GLOBAL + 3*MDRL_NUM device pointers */
  const void* const* weights_bf16; /* same indexing, tcgen05-packed bf16 matrices (may be NULL) */
  const void* const* weights_bf16_lo; /* packed residuals for GATOR_PREC_BF16X3 (may be NULL) */
  const float* pose2d;         /* (B,J,2)                                               */
  const float* pose3d;         /* (B,J,3) millimetres, or metres with pose3d_metres = 1  */
  const float* feat;           /* (B,J,128)                                             */
  float* mesh;                 /* (B,6890,3) metres                                     */
  float* coarse;               /* optional (B,431,3) coarse vertices, may be NULL        */
  void* workspace;
  size_t workspace_bytes;
} gator_mdr_args;

const char* gator_mdr_slot_name(int slot);
size_t gator_mdr_workspace_bytes(int32_t batch, int32_t num_joint, int32_t chunk);
int gator_mdr_forward(const gator_mdr_args* a, void* stream);
/* The 2-head 431x431 self-attention core (vanilla_transformer_encoder.py:36-46) on its own, fp32 FFMA kernel
 * (precision must be GATOR_PREC_FP32): qkv (B*431, 192) -> out (B*431, 64). */
int gator_mdr_self_attention(const float* qkv, float* out, int32_t batch, int32_t precision, void* stream);
/* Round-2 core of the same op: fp16 operands (fp32 accumulate), persistent warp-specialised tcgen05 kernel fed by TMA
 * bulk copies (csrc/mdr_attn2_umma.cu).  `image` is caller workspace of gator_mdr_self_attention_image_bytes(batch)
 * bytes: qkv is first re-packed into per-(sample, head) [Q | K | V] fp16 operand images (inside gator_mdr_forward the
 * layer-chain kernel writes these images itself). */
size_t gator_mdr_self_attention_image_bytes(int32_t batch);
int gator_mdr_self_attention_f16(const float* qkv, void* image, float* out, int32_t batch, void* stream);
/* The core alone on already packed operand images (what gator_mdr_forward runs; roofline measurement). */
int gator_mdr_self_attention_core(const void* image, float* out, int32_t batch, void* stream);
/* The fused row-wise chain of MDR layer `layer` (0..2) on its own (csrc/mdr_chain2_umma.cu; tensor-core precisions
 * only, always the 3-term bf16 split): everything of MDR.py:140-153 between two self-attention cores.  `weights` is the gator_mdr_args table; x_in (B*431,64) = embedded vertices (layer 0) or the
 * previous layer's x3; att_in (B*431,64) = previous self-attention output (NULL for layer 0); kv (B*J,128) = this
 * layer's cross-attention K|V; outputs x3_out (B*431,64) and q|k|v as fp32 rows qkv_out (B*431,192) and / or as the
 * fp16 operand images image_out (gator_mdr_self_attention_image_bytes(B) bytes); either of the two q|k|v outputs may
 * be NULL, not both. */
int gator_mdr_layer_chain(const void* const* weights, int32_t layer, int32_t num_joint, int32_t precision,
                          const float* x_in, const float* att_in, const float* kv, float* x3_out, float* qkv_out,
                          void* image_out, int32_t batch, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SMPL linear blend skinning - replaces SMPL_Layer.forward
 * (smplpytorch/smplpytorch/pytorch/smpl_layer.py:65-158), batch_rodrigues/quat2mat
 * (rodrigues_layer.py:13-52) and the tensutils helpers (tensutils.py:6-48).
 * ---------------------------------------------------------------------------------------------- */
#define GATOR_SMPL_JOINTS 24
#define GATOR_SMPL_K 220        /* 10 betas + 207 pose-map terms, zero-padded to a multiple of 4 */

typedef struct {
  int32_t batch;
  int32_t center_idx;          /* -1 = None                                              */
  int32_t has_betas;           /* 0: use th_betas buffer (smpl_layer.py:87-91)           */
  int32_t has_trans;           /* 0: no translation given                                */
  int32_t check_zero_norm;     /* 1: reproduce the reference's `norm(x)==0` switches on device */
  int32_t weights_per_vertex;  /* ELL width of the skinning weights                      */
  int32_t precision;
  float out_scale;             /* verts and jtr are multiplied by this after the translation; 0 = 1.0
                                  (1000.0f = the datasets' metres -> mm, data/Human36M/dataset.py:296) */
  const int32_t* parents;      /* (24) DEVICE int32 kintree parents, parents[0] ignored   */
  const float* j_template;     /* (24,3)    J_regressor @ v_template                      */
  const float* j_shapedirs;    /* (24,3,10) J_regressor @ shapedirs                       */
  const float* default_betas;  /* (10)      th_betas buffer                               */
  const float* blend_w;        /* (20670,220) [shapedirs | posedirs | 0] rows = (vertex,xyz) */
  const void* blend_w_bf16;    /* tcgen05-packed bf16 copy of blend_w, or NULL              */
  const void* blend_w_bf16_lo; /* packed residual for GATOR_PREC_BF16X3, or NULL            */
  const void* blend_w_wide;    /* tile-major split image of blend_w for the wide-N kernel (GATOR_PREC_BF16X3), or NULL */
  const float* v_template;     /* (20670)                                                 */
  const int32_t* skin_idx;     /* (6890, weights_per_vertex) joint ids                    */
  const float* skin_w;         /* (6890, weights_per_vertex)                              */
  const void* skin_w_img;      /* dense th_weights (6890,24) as the 128-row tile image of the tensor-core skinning
                                  kernel ([54][hi|lo][16][4][8][8] bf16, K padded to 32); GATOR_PREC_BF16X3, or NULL */
  const float* pose;           /* (B,72) axis-angle                                       */
  const float* betas;          /* (B,10) or NULL                                          */
  const float* trans;          /* (B,3) or NULL                                           */
  float* verts;                /* (B,6890,3)                                              */
  float* jtr;                  /* (B,24,3)                                                */
  void* workspace;
  size_t workspace_bytes;
} gator_smpl_args;

size_t gator_smpl_workspace_bytes(int32_t batch);
int gator_smpl_forward(const gator_smpl_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Any body model shaped like SMPL on the fp32 kernels - used for the MANO hand layer
 * (manopth/manopth/manolayer.py:170-230 is the arithmetic of smpl_layer.py:87-145 with 778 vertices, 16 joints and
 * 10 + 135 blend terms).  Same meaning of every field as gator_smpl_args; dimensions are arguments:
 * n_verts even, n_joints in [2, 24], k_blend = 10 + 9 (n_joints - 1) rounded up to a multiple of 4.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch;
  int32_t n_verts;
  int32_t n_joints;
  int32_t k_blend;
  int32_t center_idx;          /* -1 = None; a chain joint                                */
  int32_t has_betas;
  int32_t has_trans;
  int32_t check_zero_norm;     /* 1: `norm(x)==0` switches of smpl_layer.py:87,148 on device */
  int32_t weights_per_vertex;
  float out_scale;             /* 0 = 1.0                                                 */
  const int32_t* parents;      /* (n_joints) DEVICE int32, parents[0] ignored             */
  const float* j_template;     /* (n_joints,3)    J_regressor @ v_template                */
  const float* j_shapedirs;    /* (n_joints,3,10) J_regressor @ shapedirs                 */
  const float* default_betas;  /* (10)                                                    */
  const float* blend_w;        /* (n_verts*3, k_blend) [shapedirs | posedirs | 0]         */
  const float* v_template;     /* (n_verts*3)                                             */
  const int32_t* skin_idx;     /* (n_verts, weights_per_vertex)                           */
  const float* skin_w;         /* (n_verts, weights_per_vertex)                           */
  const float* pose;           /* (B, n_joints*3) axis-angle                              */
  const float* betas;          /* (B,10) or NULL                                          */
  const float* trans;          /* (B,3) or NULL                                           */
  float* verts;                /* (B,n_verts,3)                                           */
  float* jtr;                  /* (B,n_joints,3)                                          */
  void* workspace;
  size_t workspace_bytes;
} gator_lbs_args;

size_t gator_lbs_workspace_bytes(int32_t batch, int32_t n_verts, int32_t n_joints);
int gator_lbs_forward(const gator_lbs_args* a, void* stream);

/* MANO pose pre-step (manolayer.py:128-143): full_pose (B,48) = [coeffs[:, :3] | hands_mean + coeffs[:, 3:3+ncomps] @ comps]
 * (comps (ncomps,45); comps = NULL: use_pca = False with joint_rot_mode = 'axisang', coeffs holds the 45 axis-angle
 * values).  `ld` = row stride of coeffs in floats. */
int gator_mano_pose(const float* coeffs, int32_t ld, int32_t ncomps, const float* comps, const float* hands_mean,
                    float* full_pose, int32_t batch, void* stream);

/* MANO post-step (manolayer.py:232-256) on the outputs of gator_lbs_forward (metres, no translation): finger tips
 * sampled from the vertices, optional palm root, re-ordering into the 21-joint convention, centring on joint
 * `center_idx` of that convention or translation by `trans`, then `scale` (1000: metres -> millimetres) on vertices
 * (in place) and joints. */
typedef struct {
  int32_t batch;
  int32_t n_verts;             /* 778                                                     */
  int32_t center_idx;          /* -1 = None; index into the 21 re-ordered joints          */
  int32_t has_trans;
  int32_t check_zero_norm;     /* 1: `norm(th_trans)==0` -> centre branch, decided on the device (needs flag_ws) */
  int32_t root_palm;           /* 1: joint 0 = mean of the two palm vertices              */
  int32_t tip_verts[5];        /* 745, 317, 444 (right) | 445 (left), 556, 673            */
  int32_t palm_verts[2];       /* 95, 22                                                  */
  float scale;                 /* 0 = 1.0                                                 */
  int32_t reserved;
  const float* jtr16;          /* (B,16,3) chain joints from gator_lbs_forward            */
  const float* trans;          /* (B,3) or NULL                                           */
  float* verts;                /* (B,n_verts,3) in / out                                  */
  float* jtr;                  /* (B,21,3) out                                            */
  int32_t* flag_ws;            /* DEVICE, 4 bytes, or NULL                                */
} gator_mano_post_args;

int gator_mano_post(const gator_mano_post_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Ground-truth mesh generation, camera fix-up - the per-item host arithmetic of Human36M.get_smpl_coord
 * (data/Human36M/dataset.py:254-298) around its batch-1 SMPL forward, for a whole batch:
 *   betas'      = 0 if any |beta| > 3 else betas                                      (:266)
 *   pose'[0:3]  = axis-angle of R_cam . rodrigues(pose[0:3])  (transforms3d axangle2mat / mat2axangle, :268-274)
 *   smpl_trans  = R_cam . trans + t_cam/1000 - root + R_cam . root                     (:289-292)
 * with root = rest position of the root joint for betas' (what smpl_joint_coord[root] is when no translation
 * is passed).  Feeding pose', betas', smpl_trans and out_scale = 1000 to gator_smpl_forward gives the
 * (mesh_cam, joint_cam) of get_smpl_coord in millimetres with no pass over the mesh besides the skinning.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch;
  int32_t reserved;
  const float* j_template;     /* (24,3)    as in gator_smpl_args                          */
  const float* j_shapedirs;    /* (24,3,10)                                               */
  const float* default_betas;  /* (10) used when betas' is all zero (smpl_layer.py:87-91)  */
  const float* pose;           /* (B,72) axis-angle, world frame                          */
  const float* betas;          /* (B,10)                                                  */
  const float* trans;          /* (B,3) SMPL -> world translation, metres                 */
  const float* cam_R;          /* (B,9) row-major world -> camera rotation                */
  const float* cam_t;          /* (B,3) world -> camera translation, millimetres          */
  float* pose_out;             /* (B,72)                                                  */
  float* betas_out;            /* (B,10)                                                  */
  float* trans_out;            /* (B,3) metres                                            */
} gator_smpl_cam_args;

int gator_smpl_cam_fixup(const gator_smpl_cam_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Sparse resampling / regression - replaces spmm (lib/models/backbones/graph_layers.py:105-124) as
 * used by Mesh.downsample/upsample (lib/models/backbones/mesh.py:93-123) and the J-regression
 * post-step of every caller (lib/core/base.py:221, demo/run.py:142).
 *   y[b, r, :] = scale * sum_k val[k] * x[b, col[k], :],  k in [rowptr[r], rowptr[r+1])
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch;               /* B (1 for a 2-D input)                                   */
  int32_t rows;                /* output vertices per sample                              */
  int32_t cols;                /* input vertices per sample                               */
  int32_t feat;                /* trailing dimension (3 for xyz)                          */
  float scale;                 /* 1.0f, or 1000.0f for the metres->millimetres of base.py:219 */
  int32_t reserved;
  const int32_t* rowptr;       /* (rows+1)                                                */
  const int32_t* colidx;       /* (nnz)                                                   */
  const float* values;         /* (nnz)                                                   */
  const float* x;              /* (B, cols, feat)                                         */
  float* y;                    /* (B, rows, feat)                                         */
} gator_csr_args;

int gator_csr_spmm(const gator_csr_args* a, void* stream);

/* Two consecutive upsampling operators in one launch - Mesh.upsample(x, n1=2, n2=0) of
 * lib/models/backbones/mesh.py:110-123 (431 -> 1723 -> 6890 vertices): the intermediate level stays in shared
 * memory, so a sample costs its 5 KB input and its 83 KB output of HBM traffic and nothing else.
 * Operators are passed in ELL form as records of FOUR entries per row: col[r * 4 + k], val[r * 4 + k], k < 4
 * (16-byte aligned arrays); a row with fewer non-zeros is padded with val = 0 and a column the row already uses;
 * `width` = the largest number of non-zeros of a row (<= 4).  Arithmetic per output element is the same fmaf chain,
 * in the same order, as gator_csr_spmm applied twice.
 *   mid[b, r, :] = sum_k val1[r, k] * x[b, col1[r, k], :]            r < rows1
 *   y[b, r, :]   = scale * sum_k val2[r, k] * mid[b, col2[r, k], :]  r < rows2
 * Requires (cols + rows1) * 12 * 4 bytes <= 112 KB of shared memory (four samples per CTA, two CTAs per SM);
 * larger levels go through gator_csr_spmm twice. */
typedef struct {
  int32_t batch;               /* B                                                       */
  int32_t cols;                /* input vertices per sample (431)                         */
  int32_t rows1;               /* intermediate vertices (1723)                            */
  int32_t rows2;               /* output vertices (6890)                                  */
  int32_t width1;              /* ELL width of the first operator (1..4)                  */
  int32_t width2;              /* ELL width of the second operator (1..4)                 */
  float scale;
  int32_t reserved;
  const int32_t* col1;         /* (rows1, 4)                                              */
  const float* val1;           /* (rows1, 4)                                              */
  const int32_t* col2;         /* (rows2, 4)                                              */
  const float* val2;           /* (rows2, 4)                                              */
  const float* x;              /* (B, cols, 3)                                            */
  float* y;                    /* (B, rows2, 3)                                           */
} gator_upsample2_args;

int gator_mesh_upsample2(const gator_upsample2_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused evaluation epilogue - replaces, per batch, lib/core/base.py:219-223
 *     pred_mesh, gt_mesh = pred_mesh * 1000, gt_mesh * 1000
 *     pred_pose = J_regressor[None] @ pred_mesh
 *     j_err, s_err = val_dataset.compute_both_err(pred_mesh, gt_mesh, pred_pose, gt_pose3d)
 * with compute_both_err of data/Human36M/dataset.py:466-478 (= data/PW3D/dataset.py:273-286): root-align
 * meshes and joints on joint `root`, mean L2 over the vertices / over the evaluation joints; and the
 * per-sample MPJPE / PA-MPJPE of evaluate_joint (dataset.py:480-504, rigid_align lib/coord_utils.py:127-149).
 * One read of the two meshes, no device->host copy of a mesh.  All per-sample outputs are optional.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch;               /* B                                                       */
  int32_t verts;               /* 6890                                                    */
  int32_t joints;              /* rows of the regressor (<= 32)                           */
  int32_t n_eval;              /* evaluation joints (<= 32)                               */
  int32_t root;                /* joint both sides are aligned on (0 in both datasets)    */
  float scale;                 /* 1000.0f: metres -> millimetres applied to both meshes   */
  const int32_t* jreg_rowptr;  /* CSR J_regressor (joints x verts); unused if pred_joints_in */
  const int32_t* jreg_colidx;
  const float* jreg_values;
  const int32_t* eval_joints;  /* (n_eval) indices into the joint set                     */
  const float* pred_mesh;      /* (B, verts, 3) before `scale`                            */
  const float* gt_mesh;        /* (B, verts, 3) before `scale`, or NULL (no surface error) */
  const float* gt_joints;      /* (B, joints, 3) ALREADY in output units (reg_pose3d, mm) */
  const float* pred_joints_in; /* (B, joints, 3) in output units: skip the regression, or NULL */
  float* pred_joints;          /* out (B, joints, 3) = J_regressor @ (pred_mesh*scale), or NULL */
  float* joint_err;            /* out (B) mean over the eval joints, or NULL              */
  float* surface_err;          /* out (B) mean over the vertices, or NULL                 */
  float* pa_joint_err;         /* out (B) joint error after similarity Procrustes, or NULL (needs pred_joints) */
  float* batch_mean;           /* out (3) batch means of the three arrays above (0 where absent), or NULL */
} gator_eval_args;

int gator_eval_epilogue(const gator_eval_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 2D-pose pre-processing - replaces the per-sample numpy/cv2 chain in front of every forward:
 * add_pelvis / add_neck (demo/run.py:103-121), get_bbox + process_bbox (lib/coord_utils.py:21-66),
 * j2d_processing with rot = 0, flip = 0 (lib/aug_utils.py:51-64,140-184), "-> 0~1" and the per-axis
 * standardisation (demo/run.py:130-133, data/Human36M/dataset.py:383-389).
 * A sample whose box process_bbox would reject (returns None) gets NaN outputs and valid[b] = 0.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch;               /* B                                                       */
  int32_t joints_in;           /* joints per sample in `joints`                           */
  int32_t in_stride;           /* floats per input joint (2, or 3 with a confidence column) */
  int32_t n_mid;               /* joints to synthesise as midpoints and append (0..2)     */
  int32_t mid_a[2];            /* COCO demo: (L_Hip, R_Hip) -> Pelvis, (L_Shoulder, R_Shoulder) -> Neck */
  int32_t mid_b[2];
  float out_w;                 /* cfg.MODEL.input_shape[1] (288)                          */
  float out_h;                 /* cfg.MODEL.input_shape[0] (384)                          */
  float aspect;                /* process_bbox aspect_ratio = out_w / out_h               */
  float bbox_scale;            /* process_bbox scale (1.0)                                */
  const float* joints;         /* (B, joints_in, in_stride) pixel coordinates, x and y first */
  float* pose2d;               /* out (B, joints_in + n_mid, 2) standardised - the forward's input */
  float* joint_img;            /* out (B, joints_in + n_mid, 2) crop pixel coordinates, or NULL */
  float* bbox;                 /* out (B, 4) processed box x, y, w, h, or NULL            */
  int32_t* valid;              /* out (B), or NULL                                        */
} gator_pose2d_args;

int gator_pose2d_preprocess(const gator_pose2d_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Building block exported for tests and micro-benchmarks:
 *   C[m,n] = act( sum_k A[m,k] * W[n,k] + bias[n] + bias_rows[m % bias_period, n] ) + R[m,n]
 * A (M,K) row-major lda, W (N,K) row-major ldw (torch Linear layout); K, lda, ldw multiples of 4.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t M, N, K;
  int32_t lda, ldw, ldc, ldr;
  int32_t act;                 /* 0 none, 1 exact-erf GELU                               */
  int32_t bias_period;         /* rows of bias_rows (0 = unused)                         */
  int32_t precision;
  const float* A;
  const void* W;               /* fp32 (N,K) for GATOR_PREC_FP32; packed bf16 (see below) otherwise */
  const void* W_lo;            /* packed bf16 residual for GATOR_PREC_BF16X3, else NULL   */
  const void* W_wide;          /* GATOR_PREC_BF16X3 only: tile-major (hi|lo) image for the wide-N kernel, or NULL */
  void* a_image;               /* workspace for the split image of A when W_wide is used  */
  size_t a_image_bytes;        /* >= gator_umma_wide_a_bytes(M, K)                        */
  const float* bias;           /* (N) or NULL                                            */
  const float* bias_rows;      /* (bias_period, N) or NULL                               */
  const float* R;              /* (M, ldr) residual or NULL (may alias C)                */
  float* C;
} gator_gemm_args;

int gator_gemm(const gator_gemm_args* a, void* stream);

/* Layout of a bf16 weight packed for the tcgen05 GEMM: W (N,K) is zero-padded to (n_tiles*BN, K_pad) and
 * stored as [N_pad/8][K_pad/8][8][8] bf16 (8x8 core matrices, the UMMA K-major no-swizzle smem image). */
int gator_umma_weight_layout(int32_t N, int32_t K, int32_t* BN, int32_t* n_tiles, int32_t* K_pad);

/* Wide-N kernel (N in the thousands: upsample_conv, SMPL blend shapes).  W (N,K) is split into bf16 hi / lo and
 * stored tile-major as [n_tiles][kblocks][hi|lo][32][4][8][8] bf16: 256-row tiles, 32-wide K blocks, 8x8 core
 * matrices - one (tile, block) is one contiguous 32 KB chunk, fetched by a single TMA bulk copy. */
int gator_umma_wide_layout(int32_t N, int32_t K, int32_t* n_tiles, int32_t* kblocks);
size_t gator_umma_wide_a_bytes(int32_t M, int32_t K);

#ifdef __cplusplus
}
#endif
#endif  /* GATOR_B200_H_ */
