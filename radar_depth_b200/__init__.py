"""radar_depth_b200 -- Blackwell-native (sm_100a) implementation of the radar_depth encoder-decoder hot path.

Layout: ``csrc/`` CUDA kernels + C ABI (``include/radar_depth_b200.h``); ``_lib.py`` ctypes binding;
``convplan.py`` host-side planning of the tcgen05 convolution programs; ``engine.py`` the forward/backward
schedule of ResNet_latefusion; ``model/`` and ``evaluation/`` mirror the reference's module paths
(model/models.py, model/multistage_model.py, evaluation/criteria_new.py).
"""
__version__ = "0.1.0"
