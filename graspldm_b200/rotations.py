"""tmrp_to_H (R/grasp_ldm/utils/rotations.py:298-302) on the GPU post-processing kernel."""
import torch

from . import engine


def tmrp_to_H(tmrp):
    """[..., 6] (translation, modified Rodrigues parameters) -> [..., 4, 4] homogeneous transforms."""
    shp = tmrp.shape[:-1]
    flat = tmrp.reshape(-1, 6)
    dev = flat.device
    _, H, _ = engine.pose_postprocess(flat, torch.zeros(flat.shape[0], 1, device=dev), torch.zeros(6, device=dev),
                                      torch.ones(6, device=dev))
    return H.view(*shp, 4, 4)
