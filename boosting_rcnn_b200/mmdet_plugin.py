"""Importing this module registers the B200 classes into a real mmdet install
(use it from a config: ``custom_imports = dict(imports=
['boosting_rcnn_b200.mmdet_plugin'], allow_failed_imports=False)``)."""
from .registry import register_into_mmdet

REGISTERED = register_into_mmdet(force=True)
