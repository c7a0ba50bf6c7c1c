"""keypoint_diffusion_b200 -- B200-native sampling hot path of keypoint-diffusion.

Drop-in for the reference's sampling API (KeypointDiffusion / LigRecDynamics /
LigRecDynamicsGVP / model_from_config); all compute runs in libkpdiff_b200.so (sm_100a).
Importing the package loads the shared library and raises if it has not been built.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)
from .dynamics import LigRecDynamics, LigRecDynamicsGVP
from .hetero import HeteroBatch, build_initial_complex_graph
from .ligand_diffuser import FixedReceptorEncoder, KeypointDiffusion
from .model_setup import load_model, model_from_config
from .n_nodes_dist import LigandSizeDistribution
from .receptor_encoder import ReceptorEncoder, ReceptorEncoderGVP
from .schedule import PredefinedNoiseSchedule

__all__ = ["KeypointDiffusion", "LigRecDynamics", "LigRecDynamicsGVP", "HeteroBatch", "model_from_config",
           "load_model", "LigandSizeDistribution", "PredefinedNoiseSchedule", "ReceptorEncoder", "ReceptorEncoderGVP",
           "FixedReceptorEncoder", "build_initial_complex_graph"]
