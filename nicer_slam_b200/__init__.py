"""nicer_slam_b200 — B200-native (sm_100a) implementation of NICER-SLAM's per-iteration
neural volume-rendering hot path behind the reference's Python module API.

Plug-in: point the trainer's conf strings at this package (utils.general.get_class,
/root/reference/code/utils/general.py:153-159):
    train.model_class = "nicer_slam_b200.model.network.SLAMNetwork"
    train.loss_class  = "nicer_slam_b200.model.loss.SLAMLoss"
See INTEGRATION.md.
"""
__version__ = "0.1.0"
