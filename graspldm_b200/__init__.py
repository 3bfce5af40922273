"""graspldm_b200 - B200-native (sm_100a) implementation of GraspLDM's grasp-generation hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all arithmetic of the path
runs in hand-written CUDA kernels reached through the C ABI declared in include/graspldm_b200.h
(libgraspldm_b200.so, loaded by graspldm_b200._lib).  There is no CPU fallback: every operator
raises if the library is missing or a tensor is not on a CUDA device.

Module map (reference file each one mirrors, R = /root/reference):
  _pvcnn_backend   R/grasp_ldm/models/modules/ext/pvcnn/modules/functional/src/bindings.cpp
  functional       R/grasp_ldm/models/modules/ext/pvcnn/modules/functional/*.py
  pvcnn            R/grasp_ldm/models/modules/ext/pvcnn/{modules/*.py,pvcnn_base.py,pointnet2.py}
  pc_encoders      R/grasp_ldm/models/modules/pc_encoders.py
  resnets          R/grasp_ldm/models/modules/resnets.py
  gaussian_diffusion R/grasp_ldm/models/diffusion/gaussian_diffusion.py
  grasp_vae        R/grasp_ldm/models/grasp_vae.py
  grasp_ldm        R/grasp_ldm/models/grasp_ldm.py
  rotations        R/grasp_ldm/utils/rotations.py  (tmrp_to_H only)
  inference        R/tools/inference.py (generate_grasps of InferenceLDM / InferenceVAE)
  sharding         object-sharded multi-GPU generation with one final gather
"""
__version__ = "0.1.0"
