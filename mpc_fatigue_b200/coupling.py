"""Coupled fatigue of the dual-arm box-carrying models (config C3).

The reference's dual-arm OCPs hold one box between two arms with F_L,z + F_R,z = m g, m = 30 kg
(python/2_pilz_6_DOF/Box_Pilz_6DOF2.py:197,272-274).  In the dynamics + fatigue mode of this library the split of that load
between the arms couples their fatigue states (include/mpcf.h: mpcf_model_set_coupling; csrc/kernels_couple.cu)."""
from __future__ import annotations

from .model import Model

BOX_MASS = 30.0  # Box_Pilz_6DOF2.py:197
GRAVITY = 9.81


def box_load_coupling(model: Model, mass: float = BOX_MASS, frames: tuple[str, str] | None = None):
    """(ee_frame_arm0, ee_frame_arm1, weight) for `Model.set_coupling`: by default the frames `end_effector` /
    `sec_end_effector` of urdf/2_pilz_robot_6DOF.urdf, else the last frame carried by each arm's last joint."""
    names = model.frame_names
    if frames is None:
        if "end_effector" in names and "sec_end_effector" in names:
            frames = ("end_effector", "sec_end_effector")
        else:
            fpar = model.export("fparent").tolist()
            half = model.n // 2
            pick = lambda j: max(i for i, p in enumerate(fpar) if p == j)
            return pick(half - 1), pick(model.n - 1), mass * GRAVITY
    return model.frame_id(frames[0]), model.frame_id(frames[1]), mass * GRAVITY
