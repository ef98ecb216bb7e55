"""
gpy_dla_detection_b200 : B200-native implementation of the per-spectrum Bayesian
model-selection hot path of gpy_dla_detection (null / subDLA / multi-DLA low-rank GP
evidences over quasi-Monte-Carlo samples).

The public modules mirror the reference's: `voigt`, `effective_optical_depth`,
`set_parameters`, `null_gp`, `dla_gp`, `subdla_gp`, `dla_samples`, `subdla_samples`,
`bayesian_model_selection`, `run_bayes_select`.  All arithmetic on the path runs in
hand-written sm_100a CUDA kernels behind the C-ABI of include/dla_b200.h.
"""
__version__ = "0.1.0"
