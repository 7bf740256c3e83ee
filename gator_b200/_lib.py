"""ctypes binding of libgator_b200.so (the C ABI declared in include/gator_b200.h).

No fallback: if the shared library is missing or a call fails, a RuntimeError is raised - the product
never routes through a CPU or PyTorch-eager implementation.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, '_C', 'libgator_b200.so')

PREC_FP32, PREC_BF16, PREC_BF16X3 = 0, 1, 2
PRECISIONS = {'fp32': PREC_FP32, 'bf16': PREC_BF16, 'bf16x3': PREC_BF16X3}

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int32)


class GatArgs(C.Structure):
    _fields_ = [('num_joint', C.c_int32), ('depth', C.c_int32), ('batch', C.c_int32), ('chunk', C.c_int32),
                ('precision', C.c_int32), ('reserved', C.c_int32),
                ('weights', C.POINTER(C.c_void_p)), ('weights_bf16', C.POINTER(C.c_void_p)),
                ('weights_bf16_lo', C.POINTER(C.c_void_p)),
                ('pose2d', C.c_void_p), ('pose3d', C.c_void_p),
                ('feat', C.c_void_p), ('workspace', C.c_void_p), ('workspace_bytes', C.c_size_t)]


class MdrArgs(C.Structure):
    _fields_ = [('num_joint', C.c_int32), ('batch', C.c_int32), ('chunk', C.c_int32), ('alpha', C.c_int32),
                ('precision', C.c_int32), ('reserved', C.c_int32), ('pose3d_metres', C.c_int32), ('reserved2', C.c_int32),
                ('weights', C.POINTER(C.c_void_p)), ('weights_bf16', C.POINTER(C.c_void_p)),
                ('weights_bf16_lo', C.POINTER(C.c_void_p)),
                ('pose2d', C.c_void_p), ('pose3d', C.c_void_p),
                ('feat', C.c_void_p), ('mesh', C.c_void_p), ('coarse', C.c_void_p),
                ('workspace', C.c_void_p), ('workspace_bytes', C.c_size_t)]


class SmplArgs(C.Structure):
    _fields_ = [('batch', C.c_int32), ('center_idx', C.c_int32), ('has_betas', C.c_int32), ('has_trans', C.c_int32),
                ('check_zero_norm', C.c_int32), ('weights_per_vertex', C.c_int32), ('precision', C.c_int32),
                ('out_scale', C.c_float),
                ('parents', C.c_void_p), ('j_template', C.c_void_p), ('j_shapedirs', C.c_void_p),
                ('default_betas', C.c_void_p), ('blend_w', C.c_void_p), ('blend_w_bf16', C.c_void_p), ('blend_w_bf16_lo', C.c_void_p),
                ('blend_w_wide', C.c_void_p),
                ('v_template', C.c_void_p),
                ('skin_idx', C.c_void_p), ('skin_w', C.c_void_p), ('skin_w_img', C.c_void_p), ('pose', C.c_void_p), ('betas', C.c_void_p),
                ('trans', C.c_void_p), ('verts', C.c_void_p), ('jtr', C.c_void_p),
                ('workspace', C.c_void_p), ('workspace_bytes', C.c_size_t)]


class CsrArgs(C.Structure):
    _fields_ = [('batch', C.c_int32), ('rows', C.c_int32), ('cols', C.c_int32), ('feat', C.c_int32),
                ('scale', C.c_float), ('reserved', C.c_int32),
                ('rowptr', C.c_void_p), ('colidx', C.c_void_p), ('values', C.c_void_p),
                ('x', C.c_void_p), ('y', C.c_void_p)]


class Upsample2Args(C.Structure):
    _fields_ = [('batch', C.c_int32), ('cols', C.c_int32), ('rows1', C.c_int32), ('rows2', C.c_int32),
                ('width1', C.c_int32), ('width2', C.c_int32), ('scale', C.c_float), ('reserved', C.c_int32),
                ('col1', C.c_void_p), ('val1', C.c_void_p), ('col2', C.c_void_p), ('val2', C.c_void_p),
                ('x', C.c_void_p), ('y', C.c_void_p)]


class LbsArgs(C.Structure):
    _fields_ = [('batch', C.c_int32), ('n_verts', C.c_int32), ('n_joints', C.c_int32), ('k_blend', C.c_int32),
                ('center_idx', C.c_int32), ('has_betas', C.c_int32), ('has_trans', C.c_int32), ('check_zero_norm', C.c_int32),
                ('weights_per_vertex', C.c_int32), ('out_scale', C.c_float),
                ('parents', C.c_void_p), ('j_template', C.c_void_p), ('j_shapedirs', C.c_void_p), ('default_betas', C.c_void_p),
                ('blend_w', C.c_void_p), ('v_template', C.c_void_p), ('skin_idx', C.c_void_p), ('skin_w', C.c_void_p),
                ('pose', C.c_void_p), ('betas', C.c_void_p), ('trans', C.c_void_p), ('verts', C.c_void_p), ('jtr', C.c_void_p),
                ('workspace', C.c_void_p), ('workspace_bytes', C.c_size_t)]


class ManoPostArgs(C.Structure):
    _fields_ = [('batch', C.c_int32), ('n_verts', C.c_int32), ('center_idx', C.c_int32), ('has_trans', C.c_int32),
                ('check_zero_norm', C.c_int32), ('root_palm', C.c_int32), ('tip_verts', C.c_int32 * 5),
                ('palm_verts', C.c_int32 * 2), ('scale', C.c_float), ('reserved', C.c_int32),
                ('jtr16', C.c_void_p), ('trans', C.c_void_p), ('verts', C.c_void_p), ('jtr', C.c_void_p), ('flag_ws', C.c_void_p)]


class GemmArgs(C.Structure):
    _fields_ = [('M', C.c_int32), ('N', C.c_int32), ('K', C.c_int32),
                ('lda', C.c_int32), ('ldw', C.c_int32), ('ldc', C.c_int32), ('ldr', C.c_int32),
                ('act', C.c_int32), ('bias_period', C.c_int32), ('precision', C.c_int32),
                ('A', C.c_void_p), ('W', C.c_void_p), ('W_lo', C.c_void_p), ('W_wide', C.c_void_p), ('a_image', C.c_void_p),
                ('a_image_bytes', C.c_size_t), ('bias', C.c_void_p), ('bias_rows', C.c_void_p),
                ('R', C.c_void_p), ('C', C.c_void_p)]


class EvalArgs(C.Structure):
    _fields_ = [('batch', C.c_int32), ('verts', C.c_int32), ('joints', C.c_int32), ('n_eval', C.c_int32),
                ('root', C.c_int32), ('scale', C.c_float),
                ('jreg_rowptr', C.c_void_p), ('jreg_colidx', C.c_void_p), ('jreg_values', C.c_void_p),
                ('eval_joints', C.c_void_p), ('pred_mesh', C.c_void_p), ('gt_mesh', C.c_void_p),
                ('gt_joints', C.c_void_p), ('pred_joints_in', C.c_void_p), ('pred_joints', C.c_void_p),
                ('joint_err', C.c_void_p), ('surface_err', C.c_void_p), ('pa_joint_err', C.c_void_p),
                ('batch_mean', C.c_void_p)]


class Pose2dArgs(C.Structure):
    _fields_ = [('batch', C.c_int32), ('joints_in', C.c_int32), ('in_stride', C.c_int32), ('n_mid', C.c_int32),
                ('mid_a', C.c_int32 * 2), ('mid_b', C.c_int32 * 2),
                ('out_w', C.c_float), ('out_h', C.c_float), ('aspect', C.c_float), ('bbox_scale', C.c_float),
                ('joints', C.c_void_p), ('pose2d', C.c_void_p), ('joint_img', C.c_void_p), ('bbox', C.c_void_p),
                ('valid', C.c_void_p)]


class SmplCamArgs(C.Structure):
    _fields_ = [('batch', C.c_int32), ('reserved', C.c_int32),
                ('j_template', C.c_void_p), ('j_shapedirs', C.c_void_p), ('default_betas', C.c_void_p),
                ('pose', C.c_void_p), ('betas', C.c_void_p), ('trans', C.c_void_p), ('cam_R', C.c_void_p),
                ('cam_t', C.c_void_p), ('pose_out', C.c_void_p), ('betas_out', C.c_void_p), ('trans_out', C.c_void_p)]


_STRUCTS = [GatArgs, MdrArgs, SmplArgs, CsrArgs, GemmArgs, EvalArgs, Pose2dArgs, SmplCamArgs, Upsample2Args, LbsArgs, ManoPostArgs]
EXPORTS = ['gator_abi_version', 'gator_last_error', 'gator_abi_sizeof', 'gator_launch_count',
           'gator_mdr_self_attention', 'gator_mdr_self_attention_image_bytes', 'gator_mdr_self_attention_f16', 'gator_mdr_self_attention_core', 'gator_mdr_layer_chain', 'gator_umma_weight_layout',
           'gator_gat_slot_name', 'gator_gat_workspace_bytes', 'gator_gat_forward',
           'gator_mdr_slot_name', 'gator_mdr_workspace_bytes', 'gator_mdr_forward',
           'gator_smpl_workspace_bytes', 'gator_smpl_forward', 'gator_csr_spmm', 'gator_mesh_upsample2', 'gator_gemm',
           'gator_lbs_workspace_bytes', 'gator_lbs_forward', 'gator_mano_pose', 'gator_mano_post',
           'gator_eval_epilogue', 'gator_pose2d_preprocess', 'gator_smpl_cam_fixup',
           'gator_umma_wide_layout', 'gator_umma_wide_a_bytes']

_lock = threading.Lock()
_lib = None


def lib():
    """The loaded library; raises RuntimeError (never falls back) when it is absent or inconsistent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'gator_b200: CUDA library {LIB_PATH} is missing - run `python -m gator_b200.build` '
                '(or __graft_entry__.build()); there is no CPU fallback.')
        L = C.CDLL(LIB_PATH)
        L.gator_abi_version.restype = C.c_int
        L.gator_last_error.restype = C.c_char_p
        L.gator_abi_sizeof.restype = C.c_size_t
        L.gator_abi_sizeof.argtypes = [C.c_int]
        for name in ('gator_gat_slot_name', 'gator_mdr_slot_name'):
            getattr(L, name).restype = C.c_char_p
            getattr(L, name).argtypes = [C.c_int]
        for name in ('gator_gat_workspace_bytes', 'gator_mdr_workspace_bytes'):
            getattr(L, name).restype = C.c_size_t
            getattr(L, name).argtypes = [C.c_int32, C.c_int32, C.c_int32]
        L.gator_smpl_workspace_bytes.restype = C.c_size_t
        L.gator_smpl_workspace_bytes.argtypes = [C.c_int32]
        for name, st in (('gator_gat_forward', GatArgs), ('gator_mdr_forward', MdrArgs),
                         ('gator_smpl_forward', SmplArgs), ('gator_csr_spmm', CsrArgs), ('gator_gemm', GemmArgs),
                         ('gator_eval_epilogue', EvalArgs), ('gator_pose2d_preprocess', Pose2dArgs),
                         ('gator_smpl_cam_fixup', SmplCamArgs), ('gator_mesh_upsample2', Upsample2Args),
                         ('gator_lbs_forward', LbsArgs), ('gator_mano_post', ManoPostArgs)):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = [C.POINTER(st), C.c_void_p]
        L.gator_lbs_workspace_bytes.restype = C.c_size_t
        L.gator_lbs_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        L.gator_mano_pose.restype = C.c_int
        L.gator_mano_pose.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.gator_launch_count.restype = C.c_longlong
        L.gator_launch_count.argtypes = [C.c_int]
        L.gator_mdr_self_attention.restype = C.c_int
        L.gator_mdr_self_attention.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.gator_mdr_self_attention_image_bytes.restype = C.c_size_t
        L.gator_mdr_self_attention_image_bytes.argtypes = [C.c_int32]
        L.gator_mdr_self_attention_f16.restype = C.c_int
        L.gator_mdr_self_attention_f16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.gator_mdr_layer_chain.restype = C.c_int
        L.gator_mdr_layer_chain.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.gator_mdr_self_attention_core.restype = C.c_int
        L.gator_mdr_self_attention_core.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.gator_umma_wide_layout.restype = C.c_int
        L.gator_umma_wide_layout.argtypes = [C.c_int32, C.c_int32, c_int_p, c_int_p]
        L.gator_umma_wide_a_bytes.restype = C.c_size_t
        L.gator_umma_wide_a_bytes.argtypes = [C.c_int32, C.c_int32]
        L.gator_umma_weight_layout.restype = C.c_int
        L.gator_umma_weight_layout.argtypes = [C.c_int32, C.c_int32, c_int_p, c_int_p, c_int_p]
        if L.gator_abi_version() != 2:
            raise RuntimeError('gator_b200: ABI version mismatch between _lib.py and libgator_b200.so')
        for i, st in enumerate(_STRUCTS):
            if L.gator_abi_sizeof(i) != C.sizeof(st):
                raise RuntimeError(f'gator_b200: struct {st.__name__} is {C.sizeof(st)} B in Python, '
                                   f'{L.gator_abi_sizeof(i)} B in the library')
        _lib = L
    return _lib


def check(status: int, what: str):
    if status != 0:
        raise RuntimeError(f'{what} failed ({status}): {lib().gator_last_error().decode()}')


def slot_names(kind: str):
    """(global slot names, per-block/layer slot names) as exported by the library."""
    fn = lib().gator_gat_slot_name if kind == 'gat' else lib().gator_mdr_slot_name
    names, i = [], 0
    while True:
        s = fn(i)
        if s is None:
            break
        names.append(s.decode())
        i += 1
    # the per-block names follow the global ones; the first block name is LN1_W / N1_W
    split = names.index('LN1_W' if kind == 'gat' else 'N1_W')
    return names[:split], names[split:]


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
