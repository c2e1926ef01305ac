"""The one constant of the reference's config/config_nuscenes.py:13-14 that the hot path reads:
``PROJECT_ROOT`` (ResNet_multistage looks for ``pretrained/resnet18_latefusion.pth.tar`` under it,
model/multistage_model.py:34-39).  Set it (or the RADAR_DEPTH_PROJECT_ROOT environment variable) before
constructing ResNet_multistage(pretrained=True)."""
import os


class config_nuscenes(object):
    PROJECT_ROOT = os.environ.get("RADAR_DEPTH_PROJECT_ROOT", "YOUR_PATH/radar_depth")
