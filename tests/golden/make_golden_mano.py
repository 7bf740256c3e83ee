"""Golden vectors for the MANO layer: outputs of the UNMODIFIED reference
manopth.manolayer.ManoLayer.forward (manopth/manopth/manolayer.py:109-273) on synthetic MANO-shaped buffers
(gator_b200.synthetic.mano_data), __init__ bypassed by oracle/refshim.build_mano_layer.
Run in the build container (needs /root/reference):  python tests/golden/make_golden_mano.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from gator_b200 import synthetic   # noqa: E402
from oracle import refshim         # noqa: E402


def main():
    data = synthetic.mano_data()
    out = {}
    with torch.no_grad():
        # PCA pose space (6 components), flat hand mean, betas + translation
        pose, betas, trans = synthetic.mano_inputs(5, ncomps=6)
        t = lambda a: torch.from_numpy(a)
        layer = refshim.build_mano_layer(data, ncomps=6)
        v, j = layer(t(pose), t(betas), t(trans))
        out.update({'pca/pose': pose, 'pca/betas': betas, 'pca/trans': trans, 'pca/verts': v.numpy(), 'pca/jtr': j.numpy()})
        v, j = layer(t(pose))                                            # default betas, no translation, no centre
        out.update({'pca_plain/verts': v.numpy(), 'pca_plain/jtr': j.numpy()})
        # average hand pose as the mean, centred on joint 9, palm root, shared betas
        layer = refshim.build_mano_layer(data, center_idx=9, ncomps=12, flat_hand_mean=False, side='left')
        pose12, betas12, _ = synthetic.mano_inputs(4, ncomps=12, seed=6)
        v, j = layer(t(pose12), t(betas12), root_palm=torch.Tensor([1]), share_betas=torch.Tensor([1]))
        out.update({'centre/pose': pose12, 'centre/betas': betas12, 'centre/verts': v.numpy(), 'centre/jtr': j.numpy()})
        # centre on a finger tip (joint 8 of the 21 = tip vertex 444/445), all-zero translation given -> centre branch
        layer = refshim.build_mano_layer(data, center_idx=8, ncomps=12, flat_hand_mean=False, side='left')
        v, j = layer(t(pose12), t(betas12), torch.zeros(4, 3))
        out.update({'tip/verts': v.numpy(), 'tip/jtr': j.numpy()})
        # full 45-dimensional axis-angle pose (use_pca = False)
        layer = refshim.build_mano_layer(data, use_pca=False, ncomps=45)
        pose45, betas45, trans45 = synthetic.mano_inputs(3, ncomps=45, seed=7)
        pose45[:, 3:] *= 0.3
        v, j = layer(t(pose45), t(betas45), t(trans45))
        out.update({'full/pose': pose45, 'full/betas': betas45, 'full/trans': trans45, 'full/verts': v.numpy(), 'full/jtr': j.numpy()})
    np.savez_compressed(os.path.join(HERE, 'mano.npz'), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
